// umma_gemm.cu -- tcgen05 / TMEM TF32 GEMM for the Update product and as the self-test of the
// descriptor encodings in umma.cuh:  out[m, n] = A[m, k] * B[k, n], all row-major FP32.
//
//   A tile (128 rows x 32 k)  -> shared memory, K-major, 128-byte swizzle      (UMMA operand A)
//   B tile (32 k x N)         -> shared memory, MN-major (rows of B are contiguous along n, exactly
//                                how X / W rows lie in memory), SWIZZLE_128B_BASE32B (UMMA operand B)
//   D (128 x N, FP32)         -> TMEM, read back with tcgen05.ld for the epilogue
//
// Operands are rounded to TF32 with round-to-nearest (add half an ulp, the tensor core then drops
// the low 13 bits) on their way from global memory into shared memory, which is the reference's
// cvt.rna arithmetic (hybrid_all_kernel.cu:1102-1109, 1809-1837).  Loads are register-staged
// (ld.global.v4 -> round -> st.shared) and double buffered against the asynchronous MMAs; one
// elected thread issues tcgen05.mma, completion is tracked with tcgen05.commit -> mbarrier.
// Requirements: k % 4 == 0, n % 4 == 0, 16-byte aligned rows (else the mma.sync kernel in gemm.cu
// is used).  N tile = min(n, 256) rounded up to 16.
#include "common.cuh"
#include "umma.cuh"

namespace hcspmm {

constexpr int UG_THREADS = 256;
constexpr int UG_BM = 128;
constexpr int UG_BK = 32;
constexpr int UG_STAGES = 2;

struct UmmaGemmParams {
  const float *a, *b;
  float *out;
  long long lda, ldb, ldo;
  int m, k, n, bn;   // bn: N tile (multiple of 16, <= 256)
  int *err;          // device error flag (set on a barrier timeout)
};

__device__ __forceinline__ float4 round_tf32x4(float4 v) {
  // cvt.rna.tf32.f32 on each lane (non-finite values pass through)
  v.x = __uint_as_float(f32_to_tf32(v.x));
  v.y = __uint_as_float(f32_to_tf32(v.y));
  v.z = __uint_as_float(f32_to_tf32(v.z));
  v.w = __uint_as_float(f32_to_tf32(v.w));
  return v;
}

__global__ void __launch_bounds__(UG_THREADS, 3) umma_gemm_kernel(const UmmaGemmParams p) {
  extern __shared__ __align__(1024) uint8_t ug_smem[];
  __shared__ __align__(8) uint64_t bar_empty[UG_STAGES];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m0 = blockIdx.x * UG_BM, n0 = blockIdx.y * p.bn;
  const int bn = min(p.bn, ((p.n - n0) + 15) / 16 * 16);  // MMA N of this tile
  const int natoms = (bn + 31) / 32;
  const uint32_t a_bytes = UG_BM * 128;                    // 16 KB
  const uint32_t b_lbo = 512;                              // between 32-float n-atoms
  const uint32_t b_sbo = natoms * 512;                     // between 4-row k-atoms
  const uint32_t b_bytes = (UG_BK / 4) * b_sbo;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  // 1024-byte aligned base of the dynamic region (swizzle atoms need it)
  const uint32_t smem_base = (umma::smem_u32(ug_smem) + 1023u) & ~1023u;
  uint8_t *smem_gen = ug_smem + (smem_base - umma::smem_u32(ug_smem));

  if (tid == 0) {
    for (int s = 0; s < UG_STAGES; ++s) umma::mbar_init(&bar_empty[s], 1);
    umma::mbar_init(&bar_done, 1);
    umma::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < bn) tmem_cols <<= 1;   // power of two >= the N tile: several CTAs share the SM's 512 columns
  if (wid == 0) umma::tmem_alloc(&tmem_slot, tmem_cols);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = umma::make_idesc_tf32(UG_BM, bn, /*A K-major*/ 0, /*B MN-major*/ 1);

  const int nkb = (p.k + UG_BK - 1) / UG_BK;
  const int a_pieces = UG_BM * 8;            // 16-byte pieces of the A tile
  const int b_row_pieces = bn / 4;           // 16-byte pieces per B row (bn % 4 == 0)
  const int b_pieces = UG_BK * b_row_pieces;
  constexpr int A_PER = UG_BM * 8 / UG_THREADS;              // 4
  constexpr int B_PER_MAX = UG_BK * 64 / UG_THREADS;         // 8 (bn = 256)
  float4 ra[A_PER], rb[B_PER_MAX];
  bool ok = true;

  auto load_regs = [&](int kb) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int pid = tid + i * UG_THREADS;
      const int r = pid >> 3, c = pid & 7;
      const int row = m0 + r, kk = kb * UG_BK + c * 4;
      ra[i] = (row < p.m && kk < p.k) ? round_tf32x4(ldg_f4(p.a + (long long)row * p.lda + kk))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_PER_MAX; ++i) {
      const int pid = tid + i * UG_THREADS;
      rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pid < b_pieces) {
        const int kr = pid / b_row_pieces, ch = pid % b_row_pieces;
        const int kk = kb * UG_BK + kr, col = n0 + ch * 4;
        if (kk < p.k && col < p.n) rb[i] = round_tf32x4(ldg_f4(p.b + (long long)kk * p.ldb + col));
      }
    }
    (void)a_pieces;
  };
  auto store_stage = [&](int s) {
    uint8_t *sa = smem_gen + (uint32_t)s * stage_bytes;
    uint8_t *sb = sa + a_bytes;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int pid = tid + i * UG_THREADS;
      const int r = pid >> 3, c = pid & 7;
      *reinterpret_cast<float4 *>(sa + umma::kmajor_off(r, c * 4)) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER_MAX; ++i) {
      const int pid = tid + i * UG_THREADS;
      if (pid < b_pieces) {
        const int kr = pid / b_row_pieces, ch = pid % b_row_pieces;
        *reinterpret_cast<float4 *>(sb + umma::mnmajor_chunk_off(kr, ch, b_lbo, b_sbo)) = rb[i];
      }
    }
  };

  // zero the unused tail chunks of the last (partial) n-atom once per stage buffer so that the MMA
  // never reads uninitialised shared memory (bn may end inside an atom)
  for (int s = 0; s < UG_STAGES; ++s) {
    uint8_t *sb = smem_gen + (uint32_t)s * stage_bytes + a_bytes;
    for (uint32_t o = tid * 16; o < b_bytes; o += UG_THREADS * 16)
      *reinterpret_cast<float4 *>(sb + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  load_regs(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % UG_STAGES;
    if (kb >= UG_STAGES) {
      // the MMAs that read this buffer (k-block kb - UG_STAGES) must have completed
      ok = umma::mbar_wait(&bar_empty[s], ((kb / UG_STAGES) - 1) & 1) && ok;
    }
    store_stage(s);
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t sa = smem_base + (uint32_t)s * stage_bytes, sb = sa + a_bytes;
#pragma unroll
      for (int j = 0; j < UG_BK / 8; ++j) {
        const uint64_t da = umma::make_desc_sw128(sa + j * 32, 16, 1024);
        const uint64_t db = umma::make_desc(sb + 2 * j * b_sbo, b_lbo, b_sbo, umma::LAYOUT_SW128_BASE32B);
        umma::mma_tf32_ss(tmem_d, da, db, idesc, (kb > 0 || j > 0) ? 1u : 0u);
      }
      umma::mma_commit(&bar_empty[s]);
      if (kb == nkb - 1) umma::mma_commit(&bar_done);
    }
    if (kb + 1 < nkb) load_regs(kb + 1);   // in flight while the tensor core works
  }
  ok = umma::mbar_wait(&bar_done, 0) && ok;
  umma::tc_fence_after_sync();

  // epilogue: warp w reads TMEM lanes 32*(w%4).. (its sub-partition), column half w/4
  {
    const int lq = wid & 3, half = wid >> 2;
    const int row = m0 + lq * 32 + lane;
    for (int c0 = half * 128; c0 < half * 128 + 128 && c0 < bn; c0 += 32) {
      uint32_t v[32];
      umma::tmem_ld_32x32(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, v);
      umma::tmem_ld_wait();
      if (row < p.m) {
        float *dst = p.out + (long long)row * p.ldo + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (n0 + c0 + j + 3 < p.n && ((reinterpret_cast<uintptr_t>(dst + j) & 15) == 0)) {
            *reinterpret_cast<float4 *>(dst + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                            __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (n0 + c0 + j + t < p.n) dst[j + t] = __uint_as_float(v[j + t]);
          }
        }
      }
    }
  }
  if (!ok && p.err) atomicExch(p.err, 1);
  umma::tc_fence_before_sync();
  __syncthreads();
  if (wid == 0) umma::tmem_dealloc(tmem_d, tmem_cols);
}

bool umma_gemm_supported(const float *a, int64_t lda, const float *b, int64_t ldb, int32_t k, int32_t n) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0 && (lda & 3) == 0 &&
         (ldb & 3) == 0 && (k & 3) == 0 && (n & 3) == 0 && k > 0;
}

int launch_umma_gemm(const float *a, int64_t lda, const float *b, int64_t ldb, int32_t m, int32_t k, int32_t n,
                     float *out, int64_t ldo, int *d_err, cudaStream_t stream) {
  if (m <= 0 || n <= 0) return 0;
  UmmaGemmParams p;
  p.a = a; p.b = b; p.out = out; p.lda = lda; p.ldb = ldb; p.ldo = ldo; p.m = m; p.k = k; p.n = n;
  p.bn = n >= 256 ? 256 : (n + 15) / 16 * 16;
  p.err = d_err;
  const int natoms = (p.bn + 31) / 32;
  const size_t smem = (size_t)UG_STAGES * (UG_BM * 128 + (UG_BK / 4) * natoms * 512) + 1024;
  cudaError_t err = cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) { set_error("umma_gemm attr: %s", cudaGetErrorString(err)); return (int)err; }
  dim3 grid((m + UG_BM - 1) / UG_BM, (n + p.bn - 1) / p.bn, 1);
  umma_gemm_kernel<<<grid, UG_THREADS, smem, stream>>>(p);
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("umma_gemm launch: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

}  // namespace hcspmm
