// umma_probe.cu -- stand-alone probe of the tcgen05 building blocks in csrc/umma.cuh (debug aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe scripts/debug/umma_probe.cu && /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../hc-spmm_b200/csrc/umma.cuh"
using namespace hcspmm;

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// mode: which experiment;  out: [128][32] floats;  info: misc ints
__global__ void __launch_bounds__(128, 1) probe(int mode, float *out, int *info) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t base = (umma::smem_u32(smem) + 1023u) & ~1023u;
  uint8_t *gen = smem + (base - umma::smem_u32(smem));
  uint8_t *sa = gen, *sb = gen + 16384;
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
  if (wid == 0) umma::tmem_alloc(&slot, 64);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t td = slot;
  if (tid == 0) { info[0] = (int)td; info[1] = (int)base; }
  bool ok = true;
  if (mode >= 20) {
    // tf32, A K-major SW128, B MN-major SWIZZLE_128B_BASE32B (layout type 1):
    //   B atom = 4 k-rows x 128 B (32 floats along n); 32-byte chunk index XOR row index
    //   tile layout [k-atom][n-atom][512 B]: LBO = 512 (between n-atoms), SBO = natoms * 512 (between k-atoms)
    const int N = 64, natoms = 2;
    const uint32_t lbo = 512, sbo = natoms * 512;
    const int kk0 = (mode == 20) ? -1 : (mode == 21 ? 3 : 6);
    for (int r = tid; r < 128; r += 128)
      for (int k = 0; k < 32; ++k) {
        float v = (kk0 < 0) ? 1.f : ((k == kk0) ? (float)(r + 1) : 0.f);
        *reinterpret_cast<float *>(sa + umma::kmajor_off(r, k)) = v;
      }
    for (int i = tid; i < 8 * N; i += 128) {     // k = 0..7 (one MMA), n = 0..N-1
      const int k = i / N, n = i % N;
      const float v = (kk0 < 0) ? 1.f : (float)(k * 100 + n);
      const int katom = k >> 2, kr = k & 3, natom = n >> 5;
      const uint32_t b = (uint32_t)(n & 31) * 4;
      const uint32_t off = katom * sbo + natom * lbo + kr * 128 + ((((b >> 5) ^ kr) & 3) << 5) + (b & 31);
      *reinterpret_cast<float *>(sb + off) = v;
    }
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t idesc = umma::make_idesc_tf32(128, N, 0, 1);
      info[2] = (int)idesc;
      const uint64_t da = umma::make_desc_sw128(base, 16, 1024);
      uint64_t db = umma::make_desc_sw128(base + 16384, lbo, sbo);
      db = (db & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61);     // SWIZZLE_128B_BASE32B
      umma::mma_tf32_ss(td, da, db, idesc, 0u);
      umma::mma_commit(&bar);
    }
    ok = umma::mbar_wait(&bar, 0, 1u << 22);
    umma::tc_fence_after_sync();
    __syncthreads();
    // dump columns 32..63 for mode 22, else 0..31
    uint32_t v[32];
    umma::tmem_ld_32x32(td + ((uint32_t)(wid * 32) << 16) + (mode == 22 ? 32 : 0), v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(v[j]);
    if (tid == 0) info[3] = ok ? 1 : 0;
    umma::tc_fence_before_sync();
    __syncthreads();
    if (wid == 0) umma::tmem_dealloc(td, 64);
    return;
  } else if (mode >= 10) {
    // prefill D with 7.0 so that "MMA wrote zeros" and "MMA wrote nothing" can be told apart
    for (int c = 0; c < 32; c += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(7.0f);
      tmem_st_32x8(td + ((uint32_t)(wid * 32) << 16) + c, r);
    }
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
    // uniform operands: every 4-byte word of both regions is 1.0f (tf32) or two bf16 1.0 (0x3f803f80)
    const uint32_t word = (mode == 12 || mode == 13) ? 0x3f803f80u : 0x3f800000u;
    for (int i = tid; i < (16384 + 16384) / 4; i += 128) reinterpret_cast<uint32_t *>(gen)[i] = word;
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      uint32_t idesc;
      uint64_t da, db;
      if (mode == 10) {          // tf32, A K-major, B MN-major, SW128 (the production configuration)
        idesc = umma::make_idesc_tf32(128, 32, 0, 1);
        da = umma::make_desc_sw128(base, 16, 1024); db = umma::make_desc_sw128(base + 16384, 1024, 1024);
      } else if (mode == 11) {   // tf32, both K-major, SW128
        idesc = umma::make_idesc_tf32(128, 32, 0, 0);
        da = umma::make_desc_sw128(base, 16, 1024); db = umma::make_desc_sw128(base + 16384, 16, 1024);
      } else if (mode == 12) {   // bf16 kind::f16, both K-major, SW128
        idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        da = umma::make_desc_sw128(base, 16, 1024); db = umma::make_desc_sw128(base + 16384, 16, 1024);
      } else if (mode == 13) {   // bf16, both K-major, no swizzle: core matrices 8 x 16 B, LBO = 128 B, SBO = 256 B
        idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        da = umma::make_desc_sw128(base, 128, 256) & ~((uint64_t)7 << 61);
        db = umma::make_desc_sw128(base + 16384, 128, 256) & ~((uint64_t)7 << 61);
      } else {                   // 14: tf32 production config, accumulate onto the prefilled 7.0
        idesc = umma::make_idesc_tf32(128, 32, 0, 1);
        da = umma::make_desc_sw128(base, 16, 1024); db = umma::make_desc_sw128(base + 16384, 1024, 1024);
      }
      info[2] = (int)idesc;
      if (mode == 12 || mode == 13) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(td), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
      } else {
        umma::mma_tf32_ss(td, da, db, idesc, mode == 14 ? 1u : 0u);
      }
      umma::mma_commit(&bar);
    }
    ok = umma::mbar_wait(&bar, 0, 1u << 22);
    umma::tc_fence_after_sync();
  } else if (mode == 0) {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = __float_as_uint((float)(tid * 100 + j));
    tmem_st_32x8(td + ((uint32_t)(wid * 32) << 16), r);
  } else {
    // fill A (128 x 32 k, K-major SW128) and B (32 k x 32 n, MN-major SW128: one n-atom, 4 k-groups)
    for (int r = tid; r < 128; r += 128)
      for (int k = 0; k < 32; ++k) {
        float v = 1.f;
        if (mode == 3) v = (float)r;
        if (mode == 5) v = (float)k;
        if (mode == 6) v = (k == 3) ? 1.f : 0.f;
        if (mode == 7) v = (k == 11) ? 1.f : 0.f;
        *reinterpret_cast<float *>(sa + umma::kmajor_off(r, k)) = v;
      }
    for (int i = tid; i < 32 * 32; i += 128) {
      const int k = i / 32, n = i % 32;
      float v = 1.f;
      if (mode == 4) v = (float)n;
      if (mode == 6 || mode == 7) v = (float)(k * 100 + n);
      const uint32_t off = umma::mnmajor_chunk_off(k, n / 4, 1024u, 1024u) + (n % 4) * 4;
      *reinterpret_cast<float *>(sb + off) = v;
    }
    umma::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after_sync();
      const uint32_t idesc = umma::make_idesc_tf32(128, 32, 0, 1);
      info[2] = (int)idesc;
      const int nk = (mode == 7) ? 4 : 1;   // mode 7: all four k-steps (k = 0..31)
      for (int j = 0; j < nk; ++j) {
        const uint64_t da = umma::make_desc_sw128(base + j * 32, 16, 1024);
        const uint64_t db = umma::make_desc_sw128(base + 16384 + j * 1024, 1024, 1024);
        if (j == 0) { info[4] = (int)(da & 0xffffffffu); info[5] = (int)(da >> 32); info[6] = (int)(db & 0xffffffffu); info[7] = (int)(db >> 32); }
        umma::mma_tf32_ss(td, da, db, idesc, j > 0 ? 1u : 0u);
      }
      umma::mma_commit(&bar);
    }
    ok = umma::mbar_wait(&bar, 0, 1u << 22);
    umma::tc_fence_after_sync();
  }
  __syncthreads();
  uint32_t v[32];
  umma::tmem_ld_32x32(td + ((uint32_t)(wid * 32) << 16), v);
  umma::tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(v[j]);
  if (tid == 0) info[3] = ok ? 1 : 0;
  umma::tc_fence_before_sync();
  __syncthreads();
  if (wid == 0) umma::tmem_dealloc(td, 64);
}

int main() {
  float *d_out; int *d_info;
  cudaMalloc(&d_out, 128 * 32 * 4); cudaMalloc(&d_info, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  std::vector<float> h(128 * 32); int info[16];
  for (int mode : {20, 21, 22}) {
    cudaMemset(d_out, 0xff, 128 * 32 * 4); cudaMemset(d_info, 0, 64);
    probe<<<1, 128, 40960>>>(mode, d_out, d_info);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(info, d_info, 64, cudaMemcpyDeviceToHost);
    printf("mode %d: %s tmem=0x%x smem_base=0x%x idesc=0x%x waited=%d descA=%08x:%08x descB=%08x:%08x\n", mode, cudaGetErrorString(e),
           info[0], info[1], info[2], info[3], info[5], info[4], info[7], info[6]);
    for (int r : {0, 1, 9, 33, 127}) {
      printf("  row %3d:", r);
      for (int j = 0; j < 12; ++j) printf(" %g", h[r * 32 + j]);
      printf(" ... %g\n", h[r * 32 + 31]);
    }
    if (e != cudaSuccess) break;
  }
  return 0;
}
