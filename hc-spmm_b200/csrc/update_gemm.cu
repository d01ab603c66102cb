// update_gemm.cu -- the Update product  out[m, n] = Z[m, k] * W[k, n]  of the *_fused entry points
// (reference: hybrid_all_kernel.cu:1809-1837, both operands cvt.rna to TF32, FP32 accumulate) as a
// persistent, warp-specialised TMA + tcgen05 kernel.
//
// The product is HBM-bound at every BASELINE shape (m = 2.45 M rows, k = n = 128: 2.5 GB of Z in /
// out out against 80 GFLOP), so the kernel is built to stream:
//
//   * one CTA per SM, persistent over 128-row tiles (static round-robin);
//   * warp 0, one thread:   TMA PRODUCER -- cp.async.bulk.tensor.2d boxes of 128 rows x 32 floats of Z
//                           (and n x 32 floats of W^T) land in a ring of shared-memory stages already in
//                           the K-major SWIZZLE_128B layout tcgen05.mma reads; completion by mbarrier tx count;
//   * warps 2-5:            ROUNDERS -- cvt.rna.tf32 of the landed Z box in place (the tensor core truncates
//                           FP32 operands; the reference rounds to nearest, ties away), then a proxy fence;
//   * warp 1, one thread:   MMA ISSUER -- 4 x tcgen05.mma (M128, N = n, K8) per stage into one of TWO TMEM
//                           accumulators; tcgen05.commit frees the stage / publishes the accumulator;
//   * warps 6-9:            EPILOGUE -- tcgen05.ld 32 x 32 blocks -> swizzled staging tile in shared memory ->
//                           cp.async.bulk.tensor store (TMA clips rows >= m and columns >= n), overlapping the
//                           next tile's loads and MMAs.
//   W is transposed, rounded and zero-padded ONCE per call into a K-major scratch (<= 256 KB) by a tiny
//   pre-kernel, so that both operands take the same well-trodden K-major SWIZZLE_128B path.
// SASS: UTMALDG / UTMASTG (TMA), UTCHMMA, UTCBAR, LDTM.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "umma.cuh"

namespace hcspmm {

constexpr int TG_BM = 128;          // rows per tile (the tcgen05 M)
constexpr int TG_BK = 32;           // floats per k-block = one 128-byte swizzle row
constexpr int TG_ROUND_WARPS = 4;
constexpr int TG_EPI_WARPS = 4;
constexpr int TG_THREADS = 32 * (2 + TG_ROUND_WARPS + TG_EPI_WARPS);   // 320
constexpr int TG_MAX_STAGES = 8;
constexpr uint32_t TG_A_BYTES = TG_BM * 128;                           // 16 KB per stage
constexpr uint32_t TG_STAGING = 32 * 128;                              // 4 KB: 32 rows x 32 floats

struct TmaGemmParams {
  int m, k, n;
  int bn;        // N of one tile: multiple of 16, <= 256 (W^T scratch is zero-padded to it)
  int tiles_n;   // column tiles (> 1 only when n > 256)
  int tiles;     // row tiles x column tiles
  int stages;
  int round_a;   // 1: rounder warps apply cvt.rna to the Z boxes (0: the tensor map's TF32 type does it)
  int acc_cols;  // TMEM columns per accumulator (power of two >= bn)
  int *err;
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// W [k, n] row-major -> W^T rounded to TF32, [bn_total rows (n, zero-padded)] x [k_pad] K-major
__global__ void update_gemm_wt_kernel(const float *__restrict__ w, long long ldw, int k, int n, int k_pad, int n_pad,
                                      float *__restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k_pad * n_pad) return;
  const int col = i / k_pad, kk = i - col * k_pad;   // consecutive threads walk k: coalesced stores
  float v = 0.f;
  if (col < n && kk < k) v = __uint_as_float(f32_to_tf32(__ldg(w + (long long)kk * ldw + col)));
  wt[i] = v;
}

__global__ void __launch_bounds__(TG_THREADS, 1)
update_gemm_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const __grid_constant__ CUtensorMap tm_o, const TmaGemmParams p) {
  extern __shared__ __align__(1024) uint8_t tg_smem[];
  __shared__ __align__(8) uint64_t bar_full[TG_MAX_STAGES], bar_ready[TG_MAX_STAGES], bar_empty[TG_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t b_bytes = (uint32_t)p.bn * 128u;
  const uint32_t stage_bytes = TG_A_BYTES + b_bytes;
  const uint32_t smem_base = (umma::smem_u32(tg_smem) + 1023u) & ~1023u;
  uint8_t *gen = tg_smem + (smem_base - umma::smem_u32(tg_smem));
  const uint32_t staging_base = smem_base + (uint32_t)p.stages * stage_bytes;   // 1024-aligned: bn % 16 == 0
  const int S = p.stages;
  const int nkb = (p.k + TG_BK - 1) / TG_BK;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      umma::mbar_init(&bar_full[s], 1);
      umma::mbar_init(&bar_ready[s], TG_ROUND_WARPS * 32);
      umma::mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) { umma::mbar_init(&bar_tfull[a], 1); umma::mbar_init(&bar_tempty[a], TG_EPI_WARPS * 32); }
    umma::fence_barrier_init();
    prefetch_tensormap(&tm_a);
    prefetch_tensormap(&tm_b);
    prefetch_tensormap(&tm_o);
  }
  if (wid == 1) umma::tmem_alloc(&tmem_slot, 2u * (uint32_t)p.acc_cols);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  bool ok = true;

  if (wid == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int m0 = (t / p.tiles_n) * TG_BM, n0 = (t % p.tiles_n) * p.bn;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t s = g % S, ph = (g / S) & 1u;
          ok = umma::mbar_wait(&bar_empty[s], ph ^ 1u) && ok;           // first round: free
          const uint32_t sa = smem_base + s * stage_bytes;
          mbar_expect_tx(&bar_full[s], stage_bytes);
          tma_load_2d(sa, &tm_a, kb * TG_BK, m0, &bar_full[s]);
          tma_load_2d(sa + TG_A_BYTES, &tm_b, kb * TG_BK, n0, &bar_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (wid == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_tf32(TG_BM, p.bn, /*A K-major*/ 0, /*B K-major*/ 0);
      uint64_t *ready = p.round_a ? bar_ready : bar_full;
      uint32_t g = 0, it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, use = it >> 1;
        ok = umma::mbar_wait(&bar_tempty[acc], (use & 1u) ^ 1u) && ok;  // accumulator drained (first use: free)
        umma::tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.acc_cols;
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t s = g % S, ph = (g / S) & 1u;
          ok = umma::mbar_wait(&ready[s], ph) && ok;
          umma::fence_proxy_async_smem();
          umma::tc_fence_after_sync();
          const uint32_t sa = smem_base + s * stage_bytes, sb = sa + TG_A_BYTES;
#pragma unroll
          for (int j = 0; j < TG_BK / 8; ++j) {
            const uint64_t da = umma::make_desc_sw128(sa + j * 32, 16, 1024);
            const uint64_t db = umma::make_desc_sw128(sb + j * 32, 16, 1024);
            umma::mma_tf32_ss(tmem_d, da, db, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
          umma::mma_commit(&bar_empty[s]);
        }
        umma::mma_commit(&bar_tfull[acc]);
      }
    }
    __syncwarp();
  } else if (wid < 2 + TG_ROUND_WARPS) {
    // ================= rounders: cvt.rna.tf32 of the landed Z box, in place =================
    if (p.round_a) {
      const int rt = tid - 64;
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const uint32_t s = g % S, ph = (g / S) & 1u;
          ok = umma::mbar_wait(&bar_full[s], ph) && ok;
          float4 *a4 = reinterpret_cast<float4 *>(gen + s * stage_bytes);
#pragma unroll
          for (int i = 0; i < (int)(TG_A_BYTES / 16) / (TG_ROUND_WARPS * 32); ++i) {
            float4 v = a4[rt + i * TG_ROUND_WARPS * 32];
            v.x = __uint_as_float(f32_to_tf32(v.x));
            v.y = __uint_as_float(f32_to_tf32(v.y));
            v.z = __uint_as_float(f32_to_tf32(v.z));
            v.w = __uint_as_float(f32_to_tf32(v.w));
            a4[rt + i * TG_ROUND_WARPS * 32] = v;
          }
          umma::fence_proxy_async_smem();
          mbar_arrive1(&bar_ready[s]);
        }
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> swizzled staging -> TMA store =================
    const int lq = wid & 3;                       // the TMEM lane quarter this warp may read
    const int ew = wid - (2 + TG_ROUND_WARPS);
    const uint32_t my_staging = staging_base + (uint32_t)ew * 2u * TG_STAGING;
    uint8_t *my_gen = gen + (my_staging - smem_base);
    uint32_t it = 0, sb = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int m0 = (t / p.tiles_n) * TG_BM, n0 = (t % p.tiles_n) * p.bn;
      const uint32_t acc = it & 1u, use = it >> 1;
      ok = umma::mbar_wait(&bar_tfull[acc], use & 1u) && ok;
      umma::tc_fence_after_sync();
      const bool rows_live = m0 + lq * 32 < p.m;
      for (int c0 = 0; c0 < p.bn && n0 + c0 < p.n; c0 += 32) {
        uint32_t v[32];
        umma::tmem_ld_32x32(tmem_base + acc * (uint32_t)p.acc_cols + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0, v);
        umma::tmem_ld_wait();
        if (!rows_live) continue;
        if (lane == 0) tma_store_wait_read<1>();   // the store that last read this staging buffer has drained
        __syncwarp();
        uint8_t *dst = my_gen + sb * TG_STAGING + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4 *>(dst + ((c ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                          __uint_as_float(v[4 * c + 3]));
        umma::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_o, my_staging + sb * TG_STAGING, n0 + c0, m0 + lq * 32);
          tma_store_commit();
        }
        sb ^= 1u;
      }
      umma::tc_fence_before_sync();
      mbar_arrive1(&bar_tempty[acc]);
    }
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
  }
  if (!ok && p.err) atomicExch(p.err, 1);
  umma::tc_fence_before_sync();
  __syncthreads();
  if (wid == 1) umma::tmem_dealloc(tmem_base, 2u * (uint32_t)p.acc_cols);
}

void launch_update_gemm_wt(const float *w, int64_t ldw, int k, int n, int k_pad, int n_pad, float *wt, cudaStream_t stream) {
  update_gemm_wt_kernel<<<(k_pad * n_pad + 255) / 256, 256, 0, stream>>>(w, ldw, k, n, k_pad, n_pad, wt);
}

// ---- host side -------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  }
  return fn;
}

// 2-D FP32 tensor [rows][cols], row pitch ld floats; box = box_rows x box_cols (box_cols * 4 <= 128 when swizzled).
// swizzle: 0 none, 1 = SWIZZLE_128B (16-byte atoms: K-major tcgen05 operands), 2 = SWIZZLE_128B_ATOM_32B (the
// MN-major 32-bit operand layout, tcgen05 "SWIZZLE_128B_BASE32B").  tf32_type: the TMA unit converts FP32 -> TF32.
bool make_tensor_map_2d(CUtensorMap *map, const void *base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows, int swizzle, int tf32_type) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, tf32_type ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool update_gemm_tma_supported(const float *a, int64_t lda, const float *b, int64_t ldb, const float *out, int64_t ldo,
                               int32_t m, int32_t k, int32_t n) {
  (void)b; (void)ldb;
  return m > 0 && k > 0 && n > 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
         (lda & 3) == 0 && (ldo & 3) == 0 && tensor_map_encoder() != nullptr;
}

// wt_scratch: >= update_gemm_scratch_floats(k, n) floats of device memory (stream-ordered use)
size_t update_gemm_scratch_floats(int32_t k, int32_t n) {
  const size_t k_pad = ((size_t)k + TG_BK - 1) / TG_BK * TG_BK;
  const int bn = n >= 256 ? 256 : (n + 15) / 16 * 16;
  const size_t n_pad = ((size_t)n + bn - 1) / bn * bn;
  return k_pad * n_pad;
}

int launch_update_gemm_tma(const float *a, int64_t lda, const float *b, int64_t ldb, int32_t m, int32_t k, int32_t n,
                           float *out, int64_t ldo, float *wt_scratch, int *d_err, cudaStream_t stream) {
  TmaGemmParams p;
  p.m = m; p.k = k; p.n = n;
  p.bn = n >= 256 ? 256 : (n + 15) / 16 * 16;
  p.tiles_n = (n + p.bn - 1) / p.bn;
  p.tiles = ((m + TG_BM - 1) / TG_BM) * p.tiles_n;
  p.round_a = tuning().gemm_round != 0;
  p.acc_cols = 32;
  while (p.acc_cols < p.bn) p.acc_cols <<= 1;
  p.err = d_err;
  const int k_pad = (k + TG_BK - 1) / TG_BK * TG_BK, n_pad = p.tiles_n * p.bn;
  update_gemm_wt_kernel<<<(k_pad * n_pad + 255) / 256, 256, 0, stream>>>(b, ldb, k, n, k_pad, n_pad, wt_scratch);
  const uint32_t stage_bytes = TG_A_BYTES + (uint32_t)p.bn * 128u;
  const uint32_t staging = TG_EPI_WARPS * 2 * TG_STAGING;
  int stages = (int)((232448u - 2048u - 1024u - staging) / stage_bytes);
  if (stages > TG_MAX_STAGES) stages = TG_MAX_STAGES;
  if (tuning().gemm_stages > 0 && tuning().gemm_stages < stages) stages = tuning().gemm_stages;
  if (stages < 2) { set_error("update_gemm: tile does not fit shared memory"); return HCSPMM_E_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + staging + 1024;
  CUtensorMap tm_a, tm_b, tm_o;
  if (!make_tensor_map_2d(&tm_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, TG_BK, TG_BM, 1, p.round_a ? 0 : 1) ||
      !make_tensor_map_2d(&tm_b, wt_scratch, (uint64_t)k_pad, (uint64_t)n_pad, (uint64_t)k_pad, TG_BK, (uint32_t)p.bn, 1, 0) ||
      !make_tensor_map_2d(&tm_o, out, (uint64_t)n, (uint64_t)m, (uint64_t)ldo, 32, 32, 1, 0)) {
    set_error("update_gemm: cuTensorMapEncodeTiled failed");
    return HCSPMM_E_UNSUPPORTED;
  }
  cudaError_t err = cudaFuncSetAttribute(update_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) { set_error("update_gemm attr: %s", cudaGetErrorString(err)); return (int)err; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.tiles < sms ? p.tiles : sms;
  update_gemm_tma_kernel<<<grid, TG_THREADS, smem, stream>>>(tm_a, tm_b, tm_o, p);
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("update_gemm launch: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

}  // namespace hcspmm
