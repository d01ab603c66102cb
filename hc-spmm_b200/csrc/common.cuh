// common.cuh -- shared device helpers for the sm_100a kernels of libhcspmm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hcspmm.h"

namespace hcspmm {

constexpr int BLK_H = HCSPMM_BLK_H;  // rows per window   (reference config.h:4)
constexpr int BLK_W = HCSPMM_BLK_W;  // condensed block width (reference config.h:5)
constexpr int CTA_THREADS = 256;
constexpr int CTA_WARPS = CTA_THREADS / 32;

// thread-local error text, set by the launchers in capi.cu
void set_error(const char *fmt, ...);

struct Tuning {
  int long_row;  // rows with >= long_row non-zeros are split over all warps of the CTA
  int slab;      // feature-slab width in floats, 0 = no slabbing
  int vec8;      // use 256-bit gathers when alignment allows (1) or always 128-bit (0)
  int short_row; // rows shorter than short_row * (32 / lanes-per-row) go one lane group per row
  int wpc;       // windows per CTA, 0 = automatic from the mean window population
  int umma;      // 1: tcgen05/TMEM kernels (Update GEMM, dense super-windows) where applicable
  int pad_odd;   // 1: large operands with odd width / unaligned rows run on padded copies
  int umma_gemm; // 1: the Update GEMM uses the tcgen05 kernel (default: mma.sync kernel, still faster)
  int dense_ws;  // 1: warp-specialised dense kernel (producers / MMA issuer / double-buffered TMEM)
  int occupancy3; // 1: low-degree graphs use the 3-CTAs-per-SM build of the hybrid kernel
  int balance;    // CUDA-core windows on the merge-path balanced kernel: 0 never, 1 when nnz >= 8 * rows, 2 always
  int chunk;      // rows + stored entries per item of the balanced kernel (0 = ~4 MB of gathered rows, 4096..8192)
  int pull_ctas;  // grid cap of the halo-pull kernel (0 = 148 * 8)
  int warp_split; // items whose mean row length is >= this give every warp an equal run of entries (0 = never)
  int gemm_round;  // TMA Update GEMM: 1 = rounder warps cvt.rna the Z boxes in shared memory, 0 = TF32-typed tensor map
  int gemm_stages; // TMA Update GEMM: cap on the shared-memory ring depth (0 = as many as fit)
  int pool_keep_mb; // scratch pool: megabytes of freed blocks kept across synchronisations
  int barrier_timeout_ms; // hcspmm_peer_barrier: how long a rank waits for a peer before flagging *d_err
  int dense_tma;   // 1: dense super-windows on the TMA gather4 kernel (dense_tma.cu), 0: cp.async kernels (dense.cu)
  int fuse_update; // 1: Aggregation + Update as one kernel when the dense plan covers the graph
  int dense_min_rowlen; // dense plan: minimum mean stored entries per row of a 128-row super-window
  int l2_hot_mb;   // balanced kernel with tagged column ids: megabytes of X rows kept L2-resident (evict_last); 0 = off
  int l2_hot_min_row; // ... applied only to gathers of at least this many bytes per row
  int staged;      // 1: low-degree graphs with 260..512-byte rows gather through shared-memory staging (spmm_staged_kernel)
};
Tuning &tuning();

// Stream-ordered scratch from the library's PRIVATE memory pool (one per device): freed blocks up to
// "pool_keep_mb" stay cached across synchronisations (a fresh mapping per call costs more than the
// kernels it serves); the device's default pool and PyTorch's caching allocator are left alone.
cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t stream);
void scratch_free(void *ptr, cudaStream_t stream);

// ---- small PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}

__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// D(16x8,f32) += A(16x8,tf32,row) * B(8x8,tf32,col)
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4],
                                                 uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
      "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src, int src_bytes) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src),
               "r"(src_bytes));
}
// the same through L1 (.ca): the warp's 16-byte pieces are merged into whole sectors before they go to L2 -- the L1-bypassing
// form fetches a 32-byte sector per 16-byte piece when the pieces of a warp cover a contiguous row (measured: 2 x the bytes)
__device__ __forceinline__ void cp_async_16_ca(void *smem_dst, const void *gmem_src, int src_bytes) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// a += b with two packed FP32 adds (FADD2, new on sm_100): same IEEE result as four FADDs
__device__ __forceinline__ void add4(float4 &a, const float4 &b) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  a = make_float4(lo.x, lo.y, hi.x, hi.y);
}

}  // namespace hcspmm
