// peer.cu -- the per-layer X exchange of the row-partitioned SpMM over NVLink peer memory.
//
// The reference has no multi-GPU code.  A row shard of A references only some rows of X ("halo":
// 0.31 N remote rows for a 1/8 shard of the products shape, where an all-gather moves 0.875 N), so
// instead of a collective every rank PULLS exactly the rows its shard references, straight out of
// the owners' memory with NVLink loads:
//   * every rank keeps its SpMM operand [halo rows of lower ranks | own rows | halo rows of higher ranks]
//     in a cudaMalloc'ed buffer exported with CUDA IPC (hcspmm_peer_alloc / hcspmm_peer_open) -- one
//     process per GPU, any launcher; the own rows are written in place and are what the peers read;
//   * hcspmm_peer_barrier: one tiny kernel; thread s stores this rank's epoch into peer s's flag
//     array (system-scope release) and spins on the local flag of peer s (acquire), so after it every
//     peer's shard written before ITS barrier is visible.  Spins are bounded (knob "barrier_timeout_ms",
//     default 10 s) and report through *d_err (1 + the missing peer) instead of hanging the device: the
//     result of an aggregation whose barrier timed out is undefined, and the host layer raises on it;
//   * hcspmm_halo_pull: operand row i (owner s = segment of i, row src_row[i] there) is copied with
//     128-bit loads from peer_x[s]; eight rows in flight per warp cover the NVLink latency.
// Buffers are used alternately (two per width) by the caller, so one barrier per aggregation suffices:
// a shard buffer is rewritten only after the next barrier, which every peer enters after its pull.
#include <string.h>

#include "common.cuh"

namespace hcspmm {

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_barrier_kernel(int *const *flag_ptrs, int rank, int world, int epoch, int *err,
                                    unsigned long long timeout_ns) {
  const int s = threadIdx.x;
  if (s >= world) return;
  __threadfence_system();
  int *remote = flag_ptrs[s] + rank;   // peer s's flag for me
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const int *local = flag_ptrs[rank] + s;   // my flag for peer s
  const unsigned long long t0 = global_ns();
  int v, spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
    if (v - epoch >= 0) break;
    if ((++spins & 1023) == 0 && global_ns() - t0 > timeout_ns) {
      // a peer never arrived: what the following pull reads is UNDEFINED.  *err may live in pinned host memory
      // (hcspmm.peer does that) so the host sees it without a synchronisation and raises at its next call.
      if (err) { *reinterpret_cast<volatile int *>(err) = 1 + s; __threadfence_system(); }
      break;
    }
    __nanosleep(64);
  } while (true);
  __threadfence_system();
}

constexpr int PULL_ROWS = 8;   // rows in flight per warp

// dst[i, c0 .. c0+width) = peer_x[s][src_row[i], c0 .. c0+width) for the operand rows i of every owner s
// in owner_mask (segment [seg[s], seg[s+1])).  Batches of PULL_ROWS rows are dealt round-robin over the
// owners, starting at `first`: the warps of a CTA read from different peers at the same time and the ranks
// start on different owners, so no owner's NVLink egress is the one everybody waits for.
__global__ void __launch_bounds__(256) halo_pull_kernel(const float *const *__restrict__ peer_x, long long lds,
                                                        const int *__restrict__ src_row, const int *__restrict__ seg,
                                                        int world, unsigned long long owner_mask, int first, int c0,
                                                        int width, float *__restrict__ dst, long long ldd,
                                                        const int *__restrict__ dst_row) {
  __shared__ int s_seg[65];
  __shared__ const float *s_base[64];
  __shared__ int s_list[64];
  __shared__ int s_nb, s_ninc;
  for (int i = threadIdx.x; i <= world; i += blockDim.x) s_seg[i] = seg[i];
  for (int i = threadIdx.x; i < world; i += blockDim.x) s_base[i] = peer_x[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int nb = 0, ninc = 0;
    for (int i = 0; i < world; ++i) {
      const int o = (first + i) % world;
      if (!((owner_mask >> o) & 1ull)) continue;
      s_list[ninc++] = o;
      nb = max(nb, (s_seg[o + 1] - s_seg[o] + PULL_ROWS - 1) / PULL_ROWS);
    }
    s_nb = nb;
    s_ninc = ninc;
  }
  __syncthreads();
  const int ninc = s_ninc;
  if (ninc == 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = width >> 2;
  const long long total = (long long)s_nb * ninc;
  for (long long t = warp; t < total; t += n_warps) {
    const int o = s_list[(int)(t % ninc)];
    const int r0 = s_seg[o] + (int)(t / ninc) * PULL_ROWS, r1 = s_seg[o + 1];
    if (r0 >= r1) continue;
    const float *base = s_base[o] + c0;
    const float *src[PULL_ROWS];
#pragma unroll
    for (int j = 0; j < PULL_ROWS; ++j) src[j] = base + (long long)__ldg(src_row + min(r0 + j, r1 - 1)) * lds;
    // list entry i lands in operand row i, or (row-block pipeline: the list is a SUBSET of the halo) in dst_row[i]
    float *drow[PULL_ROWS];
#pragma unroll
    for (int j = 0; j < PULL_ROWS; ++j) {
      const int i = min(r0 + j, r1 - 1);
      drow[j] = dst + (long long)(dst_row ? __ldg(dst_row + i) : i) * ldd + c0;
    }
    for (int v = lane; v < nvec; v += 32) {
      float4 x[PULL_ROWS];
#pragma unroll
      for (int j = 0; j < PULL_ROWS; ++j) x[j] = *reinterpret_cast<const float4 *>(src[j] + v * 4);
#pragma unroll
      for (int j = 0; j < PULL_ROWS; ++j)
        if (r0 + j < r1) *reinterpret_cast<float4 *>(drow[j] + v * 4) = x[j];
    }
  }
}

// The same exchange as a PUSH: the OWNER writes the rows each peer's shard references straight into that peer's
// operand buffer (NVLink stores are posted -- no request/response round trip as with loads).  send_row[j] for
// j in [send_seg[s], send_seg[s+1]) are this rank's row indices peer s wants, ascending; they land in peer s's operand
// at rows dst_base[s] + (j - send_seg[s]) (dst_base[s] already points at this rank's segment there).  Batches are dealt
// round-robin over the peers starting at `first`, as in the pull.  The caller's next hcspmm_peer_barrier publishes
// the rows (st.release.sys after a system fence) before any peer's SpMM reads them.
__global__ void __launch_bounds__(256) halo_push_kernel(const float *__restrict__ src, long long lds,
                                                        const int *__restrict__ send_row, const int *__restrict__ send_seg,
                                                        float *const *__restrict__ dst_base, long long ldd, int world,
                                                        unsigned long long peer_mask, int first, int c0, int width) {
  __shared__ int s_seg[65];
  __shared__ float *s_base[64];
  __shared__ int s_list[64];
  __shared__ int s_nb, s_ninc;
  for (int i = threadIdx.x; i <= world; i += blockDim.x) s_seg[i] = send_seg[i];
  for (int i = threadIdx.x; i < world; i += blockDim.x) s_base[i] = dst_base[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int nb = 0, ninc = 0;
    for (int i = 0; i < world; ++i) {
      const int o = (first + i) % world;
      if (!((peer_mask >> o) & 1ull)) continue;
      s_list[ninc++] = o;
      nb = max(nb, (s_seg[o + 1] - s_seg[o] + PULL_ROWS - 1) / PULL_ROWS);
    }
    s_nb = nb;
    s_ninc = ninc;
  }
  __syncthreads();
  const int ninc = s_ninc;
  if (ninc == 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = width >> 2;
  const long long total = (long long)s_nb * ninc;
  for (long long t = warp; t < total; t += n_warps) {
    const int o = s_list[(int)(t % ninc)];
    const int j0 = s_seg[o] + (int)(t / ninc) * PULL_ROWS, j1 = s_seg[o + 1];
    if (j0 >= j1) continue;
    const float *from[PULL_ROWS];
#pragma unroll
    for (int j = 0; j < PULL_ROWS; ++j) from[j] = src + (long long)__ldg(send_row + min(j0 + j, j1 - 1)) * lds + c0;
    float *to = s_base[o] + (long long)(j0 - s_seg[o]) * ldd + c0;
    for (int v = lane; v < nvec; v += 32) {
      float4 x[PULL_ROWS];
#pragma unroll
      for (int j = 0; j < PULL_ROWS; ++j) x[j] = ldg_f4(from[j] + v * 4);
#pragma unroll
      for (int j = 0; j < PULL_ROWS; ++j)
        if (j0 + j < j1) *reinterpret_cast<float4 *>(to + (long long)j * ldd + v * 4) = x[j];
    }
  }
}

}  // namespace hcspmm

using namespace hcspmm;

extern "C" {

int hcspmm_peer_alloc(size_t bytes, void **d_ptr, void *handle64) {
  if (!d_ptr || !handle64 || bytes == 0) { set_error("peer_alloc: bad argument"); return HCSPMM_E_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    set_error("peer_alloc: %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  *d_ptr = p;
  return 0;
}

int hcspmm_peer_open(const void *handle64, void **d_ptr) {
  if (!d_ptr || !handle64) { set_error("peer_open: bad argument"); return HCSPMM_E_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("peer_open: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int hcspmm_peer_close(void *d_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
  if (e != cudaSuccess) { set_error("peer_close: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int hcspmm_peer_free(void *d_ptr) {
  cudaError_t e = cudaFree(d_ptr);
  if (e != cudaSuccess) { set_error("peer_free: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int hcspmm_peer_barrier(int32_t *const *d_flag_ptrs, int32_t rank, int32_t world, int32_t epoch, int32_t *d_err,
                        void *stream) {
  if (!d_flag_ptrs || world < 1 || world > 64 || rank < 0 || rank >= world) {
    set_error("peer_barrier: bad argument");
    return HCSPMM_E_INVALID;
  }
  const unsigned long long timeout_ns =
      (unsigned long long)(tuning().barrier_timeout_ms > 0 ? tuning().barrier_timeout_ms : 10000) * 1000000ull;
  peer_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(reinterpret_cast<int *const *>(d_flag_ptrs), rank, world, epoch,
                                                        d_err, timeout_ns);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("peer_barrier: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int hcspmm_halo_pull(const float *const *d_peer_x, int64_t lds, const int32_t *d_src_row, const int32_t *d_seg,
                     int32_t world, uint64_t owner_mask, int32_t first_owner, int32_t rows, int32_t col0, int32_t width,
                     float *d_dst, int64_t ldd, void *stream) {
  return hcspmm_halo_pull_rows(d_peer_x, lds, d_src_row, nullptr, d_seg, world, owner_mask, first_owner, rows, col0, width,
                               d_dst, ldd, stream);
}

int hcspmm_halo_pull_rows(const float *const *d_peer_x, int64_t lds, const int32_t *d_src_row, const int32_t *d_dst_row,
                          const int32_t *d_seg, int32_t world, uint64_t owner_mask, int32_t first_owner, int32_t rows,
                          int32_t col0, int32_t width, float *d_dst, int64_t ldd, void *stream) {
  if (rows <= 0 || width == 0 || owner_mask == 0) return 0;
  if (!d_peer_x || !d_src_row || !d_seg || !d_dst || world < 1 || world > 64 || width < 0 || col0 < 0 ||
      first_owner < 0) {
    set_error("halo_pull: bad argument");
    return HCSPMM_E_INVALID;
  }
  if ((width & 3) || (col0 & 3) || (lds & 3) || (ldd & 3) || (reinterpret_cast<uintptr_t>(d_dst) & 15)) {
    set_error("halo_pull: width, col0 and leading dims must be multiples of 4 floats, dst 16-byte aligned");
    return HCSPMM_E_ALIGN;
  }
  const long long warps = ((long long)rows + PULL_ROWS - 1) / PULL_ROWS + world;
  long long grid = (warps + 7) / 8;
  const long long cap = tuning().pull_ctas > 0 ? tuning().pull_ctas : 148 * 8;
  if (grid > cap) grid = cap;
  halo_pull_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_peer_x, lds, d_src_row, d_seg, world, owner_mask,
                                                                     first_owner % world, col0, width, d_dst, ldd, d_dst_row);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("halo_pull: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int hcspmm_halo_push(const float *d_src, int64_t lds, const int32_t *d_send_row, const int32_t *d_send_seg,
                     float *const *d_dst_base, int64_t ldd, int32_t world, uint64_t peer_mask, int32_t first_peer,
                     int32_t rows, int32_t col0, int32_t width, void *stream) {
  if (rows <= 0 || width == 0 || peer_mask == 0) return 0;
  if (!d_src || !d_send_row || !d_send_seg || !d_dst_base || world < 1 || world > 64 || width < 0 || col0 < 0 ||
      first_peer < 0) {
    set_error("halo_push: bad argument");
    return HCSPMM_E_INVALID;
  }
  if ((width & 3) || (col0 & 3) || (lds & 3) || (ldd & 3) || (reinterpret_cast<uintptr_t>(d_src) & 15)) {
    set_error("halo_push: width, col0 and leading dims must be multiples of 4 floats, src 16-byte aligned");
    return HCSPMM_E_ALIGN;
  }
  const long long warps = ((long long)rows + PULL_ROWS - 1) / PULL_ROWS + world;
  long long grid = (warps + 7) / 8;
  const long long cap = tuning().pull_ctas > 0 ? tuning().pull_ctas : 148 * 8;
  if (grid > cap) grid = cap;
  halo_push_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_src, lds, d_send_row, d_send_seg, d_dst_base, ldd,
                                                                     world, peer_mask, first_peer % world, col0, width);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("halo_push: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // extern "C"
