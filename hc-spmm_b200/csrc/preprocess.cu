// preprocess.cu -- GPU preprocessing, bit-exact with the reference's preprocess()
// (/root/reference/hybrid_kernel/hybrid_all_kernel.cu:339-408) but window-parallel.
//
// The reference does: fill_edgeToRow (:314-326), fill_segment (:289-301), a GLOBAL
// thrust::sort of (window, column) pairs (:386-399), then generate_edgetocolumn
// (:242-269) with ONE THREAD per window doing a serial dedup and a binary search per
// edge.  What those steps compute per 16-row window is
//     U        = number of distinct column ids in the window
//     rank(c)  = number of distinct window columns smaller than c      -> edgeToColumn
//     blockPartition = ceil(U / 8),   hybrid_type = selector(U - 1, E_w, blockPartition)
// so no global sort is needed.  Here one CTA owns one window:
//   * windows with <= SORT_CAP edges: bitonic sort of (column, edge) keys in shared
//     memory, head flags + block scan give the ranks;
//   * larger windows (queued on a device work list): a shared-memory BITMAP over the
//     window's column span, popcount prefix per word, rank = prefix + popc(masked word).
// Both are O(E_w) global traffic (column ids read twice, L2-resident).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace hcspmm {

constexpr int SORT_CAP = 4096;
constexpr int SMALL_THREADS = 256;
constexpr int LARGE_THREADS = 512;

// The core selector.  Modes 0/1 reproduce, operation for operation, what nvcc emits for
// hybrid_all_kernel.cu:262 / :261 (cvt.rn.f32.s32, cvt.rn.f32.u32, div.rn.f32, mul.f64,
// fma.rn.f64, add.f64, setp.eq / setp.leu) -- see oracle/hcspmm_oracle.c.
__device__ __forceinline__ int classify(int size, unsigned n_edges, int num, int mode) {
  if (mode == HCSPMM_CLASSIFIER_ALL_CUDA) return 0;
  if (mode == HCSPMM_CLASSIFIER_ALL_TC) return 1;
  if (mode == HCSPMM_CLASSIFIER_B200) {
    // B200 re-fit (DESIGN.md "selector"): the dense path gathers each distinct column once
    // per window, so it wins as soon as the mean column reuse E_w / U covers its fixed cost.
    int upad = num * BLK_W;
    // ... and the window must fit the per-window path's residency (<= 1024 condensed columns):
    // hub windows with tens of thousands of distinct columns belong to the CUDA-core path.
    // Label 3 = "tensor-core CANDIDATE": it is computed on tcgen05 when a dense super-window plan covers it
    // and on the CUDA cores otherwise -- the per-window mma.sync path (label 1) is never faster than the
    // CUDA-core path on B200 (format sweep, profiles/README.md), only 128-row super-windows are.
    return (n_edges >= 24u && 2u * n_edges >= 3u * (unsigned)upad && upad <= 1024) ? 3 : 0;
  }
  float sf = (float)size;
  float df = __fdiv_rn((float)n_edges, (float)(int)((unsigned)num << 7));
  if (mode == HCSPMM_CLASSIFIER_B200_WINDOW) {
    // The reference's own recipe (technical report IV-C) re-run on B200 (benchmarks/selector_fit.py,
    // profiles/r2_selector_fit.json): 16-row synthetic windows, 1..130 distinct columns, sparsity 1/16..15/16, CUDA-core
    // path against the per-window mma.sync path at dim 128, logistic regression on the reference's two features.
    // Same form as :261 -- score > 0 -> CUDA cores -- with the B200 coefficients; no U <= 32 guard (this
    // implementation has no MAX_BLK), but the per-window path keeps <= 1024 condensed columns resident.
    const double zb = (double)sf * -0.02312523 + (double)df * -9.74306426 + 4.93743285;
    return (!(zb > 0.0) && num * BLK_W <= 1024) ? 1 : 0;
  }
  double t = __dmul_rn((double)df, -6.578043);
  double u = __fma_rn((double)sf, 0.19854024, t);
  double z = __dadd_rn(u, -3.14922857);
  if (mode == HCSPMM_CLASSIFIER_SHIPPED) return z == 0.0 ? 1 : 0;
  if (size > 32) return 0;
  return !(z > 0.0) ? 1 : 0;
}

template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_buf, int &total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect warp_buf reuse across calls
  if (lane == 31) warp_buf[wid] = inc;
  __syncthreads();
  int wsum = (lane < THREADS / 32) ? warp_buf[lane] : 0;
  int winc = wsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  total = __shfl_sync(0xffffffffu, winc, THREADS / 32 - 1);
  int wexc = __shfl_sync(0xffffffffu, winc - wsum, wid);
  return wexc + inc - v;
}

// row of edge e inside the window: largest r in [0,H) with rp[r] <= e   (H a power of two)
template <int H>
__device__ __forceinline__ int row_of_edge(const int *rp, int e) {
  int lo = 0, hi = H;
#pragma unroll
  for (int s = 1; s < H; s <<= 1) {
    int mid = (lo + hi) >> 1;
    if (rp[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// H = rows per window: 16 for the reference's row windows (BLK_H); 128 for the super-windows of the
// tcgen05 dense plan, where only the ranks and the distinct-column counts are used.
template <int H>
__global__ void __launch_bounds__(SMALL_THREADS)
preprocess_small_kernel(const int *__restrict__ colidx, const int *__restrict__ rowptr,
                        int n_rows, int n_windows, int mode, int *__restrict__ block_partition,
                        int *__restrict__ edge_to_column, int *__restrict__ edge_to_row,
                        int *__restrict__ hybrid_type, int *__restrict__ large_list,
                        int *__restrict__ large_count) {
  __shared__ unsigned long long keys[SORT_CAP];
  __shared__ int rp[H + 1];
  __shared__ int warp_buf[32];
  const int w = blockIdx.x, tid = threadIdx.x;
  const int r0 = w * H;
  if (tid <= H) rp[tid] = (r0 < n_rows) ? rowptr[min(r0 + tid, n_rows)] : 0;
  __syncthreads();
  const int e0 = rp[0], e1 = rp[H];
  const int ne = e1 - e0;
  if (ne <= 0) {  // reference returns before writing (:252-253); defined as 0 here
    if (tid == 0) { block_partition[w] = 0; hybrid_type[w] = 0; }
    return;
  }
  if (edge_to_row != nullptr)
    for (int e = e0 + tid; e < e1; e += SMALL_THREADS) edge_to_row[e] = r0 + row_of_edge<H>(rp, e);
  if (ne > SORT_CAP) {
    if (tid == 0) large_list[atomicAdd(large_count, 1)] = w;
    return;
  }
  int P = 2;
  while (P < ne) P <<= 1;
  for (int i = tid; i < P; i += SMALL_THREADS)
    keys[i] = i < ne ? (((unsigned long long)(unsigned)colidx[e0 + i] << 32) | (unsigned)i)
                     : ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += SMALL_THREADS) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int p = i | j;
        unsigned long long a = keys[i], b = keys[p];
        bool asc = (i & k) == 0;
        if ((a > b) == asc) { keys[i] = b; keys[p] = a; }
      }
      __syncthreads();
    }
  }
  // head flags -> ranks.  Each thread owns a contiguous run of the sorted keys.
  const int ipt = (P + SMALL_THREADS - 1) / SMALL_THREADS;
  const int b0 = tid * ipt, b1 = min(b0 + ipt, ne);
  int local = 0;
  for (int i = b0; i < b1; ++i)
    local += (i == 0 || (unsigned)(keys[i] >> 32) != (unsigned)(keys[i - 1] >> 32));
  int total;
  int run = block_exclusive_scan<SMALL_THREADS>(local, warp_buf, total);
  for (int i = b0; i < b1; ++i) {
    run += (i == 0 || (unsigned)(keys[i] >> 32) != (unsigned)(keys[i - 1] >> 32));
    edge_to_column[e0 + (int)(unsigned)(keys[i] & 0xffffffffu)] = run - 1;
  }
  if (tid == 0) {
    int size = total - 1;                 // :256-257, loc = #unique - 1
    int num = (size + BLK_W) / BLK_W;     // :258
    block_partition[w] = num;             // :260
    hybrid_type[w] = classify(size, (unsigned)ne, num, mode);
  }
}

// Persistent CTAs over the work list of windows with more than SORT_CAP edges.
// Dynamic shared memory: bitmap[chunk_words] + prefix[chunk_words].
template <int H>
__global__ void __launch_bounds__(LARGE_THREADS)
preprocess_large_kernel(const int *__restrict__ colidx, const int *__restrict__ rowptr,
                        int n_rows, int mode, int chunk_words, int *__restrict__ block_partition,
                        int *__restrict__ edge_to_column, int *__restrict__ hybrid_type,
                        const int *__restrict__ large_list, const int *__restrict__ large_count) {
  extern __shared__ unsigned smem_u[];
  unsigned *bitmap = smem_u;
  int *prefix = reinterpret_cast<int *>(smem_u + chunk_words);
  __shared__ int warp_buf[32];
  __shared__ int s_min, s_max;
  const int tid = threadIdx.x;
  const int n_large = *large_count;
  const int chunk_bits = chunk_words * 32;
  for (int li = blockIdx.x; li < n_large; li += gridDim.x) {
    const int w = large_list[li];
    const int r0 = w * H;
    const int e0 = rowptr[r0], e1 = rowptr[min(r0 + H, n_rows)];
    if (tid == 0) { s_min = 0x7fffffff; s_max = -1; }
    __syncthreads();
    int mn = 0x7fffffff, mx = -1;
    for (int e = e0 + tid; e < e1; e += LARGE_THREADS) {
      int c = colidx[e];
      mn = min(mn, c);
      mx = max(mx, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) { atomicMin(&s_min, mn); atomicMax(&s_max, mx); }
    __syncthreads();
    const int cmin = s_min & ~31, cmax = s_max;
    int base = 0;
    for (long long c0 = cmin; c0 <= cmax; c0 += chunk_bits) {
      const int nwords = (int)min((long long)chunk_words, (((long long)cmax - c0) >> 5) + 1);
      for (int i = tid; i < nwords; i += LARGE_THREADS) bitmap[i] = 0u;
      __syncthreads();
      for (int e = e0 + tid; e < e1; e += LARGE_THREADS) {
        long long d = (long long)colidx[e] - c0;
        if (d >= 0 && d < chunk_bits) atomicOr(&bitmap[d >> 5], 1u << (d & 31));
      }
      __syncthreads();
      const int wpt = (nwords + LARGE_THREADS - 1) / LARGE_THREADS;
      const int w0 = min(tid * wpt, nwords), w1 = min(w0 + wpt, nwords);
      int local = 0;
      for (int i = w0; i < w1; ++i) {
        prefix[i] = local;
        local += __popc(bitmap[i]);
      }
      int total;
      int off = block_exclusive_scan<LARGE_THREADS>(local, warp_buf, total) + base;
      for (int i = w0; i < w1; ++i) prefix[i] += off;
      __syncthreads();
      for (int e = e0 + tid; e < e1; e += LARGE_THREADS) {
        long long d = (long long)colidx[e] - c0;
        if (d >= 0 && d < chunk_bits) {
          unsigned word = bitmap[d >> 5];
          edge_to_column[e] = prefix[d >> 5] + __popc(word & ((1u << (d & 31)) - 1u));
        }
      }
      base += total;
      __syncthreads();
    }
    if (tid == 0) {
      int size = base - 1;
      int num = (size + BLK_W) / BLK_W;
      block_partition[w] = num;
      hybrid_type[w] = classify(size, (unsigned)(e1 - e0), num, mode);
    }
    __syncthreads();
  }
}

template <int H>
static int launch_preprocess_h(const int32_t *colidx, const int32_t *rowptr, int32_t n_rows, int32_t n_windows,
                               int mode, int32_t *bp, int32_t *etc, int32_t *etr, int32_t *ht, void *ws,
                               cudaStream_t stream);

size_t preprocess_workspace_bytes(int32_t n_rows, int64_t /*nnz*/) {
  size_t w = ((size_t)n_rows + BLK_H - 1) / BLK_H;
  return (w + 64) * sizeof(int);  // [0] = work-list counter, [16..] = work list
}

int launch_preprocess(const int32_t *colidx, const int32_t *rowptr, int32_t n_rows, int64_t nnz,
                      int32_t n_windows, int mode, int32_t *bp, int32_t *etc, int32_t *etr,
                      int32_t *ht, void *ws, size_t ws_bytes, cudaStream_t stream) {
  if (n_rows < 0 || nnz < 0 || n_windows != (n_rows + BLK_H - 1) / BLK_H) {
    set_error("preprocess: n_windows must be ceil(n_rows/16) (got %d for %d rows)", n_windows,
              n_rows);
    return HCSPMM_E_INVALID;
  }
  if (mode < 0 || mode > HCSPMM_CLASSIFIER_B200_WINDOW) {
    set_error("preprocess: unknown classifier mode %d", mode);
    return HCSPMM_E_INVALID;
  }
  if (n_windows == 0) return 0;
  if (!colidx && nnz > 0) { set_error("preprocess: null colidx"); return HCSPMM_E_INVALID; }
  if (!rowptr || !bp || !ht || !ws || (nnz > 0 && (!etc || !etr))) {
    set_error("preprocess: null pointer argument");
    return HCSPMM_E_INVALID;
  }
  if (ws_bytes < preprocess_workspace_bytes(n_rows, nnz)) {
    set_error("preprocess: workspace too small (%zu < %zu)", ws_bytes,
              preprocess_workspace_bytes(n_rows, nnz));
    return HCSPMM_E_WORKSPACE;
  }
  return launch_preprocess_h<BLK_H>(colidx, rowptr, n_rows, n_windows, mode, bp, etc, etr, ht, ws, stream);
}

template <int H>
static int launch_preprocess_h(const int32_t *colidx, const int32_t *rowptr, int32_t n_rows, int32_t n_windows,
                               int mode, int32_t *bp, int32_t *etc, int32_t *etr, int32_t *ht, void *ws,
                               cudaStream_t stream) {
  int *counter = reinterpret_cast<int *>(ws);
  int *list = counter + 16;
  cudaError_t err = cudaMemsetAsync(counter, 0, 16 * sizeof(int), stream);
  if (err != cudaSuccess) { set_error("preprocess: memset: %s", cudaGetErrorString(err)); return (int)err; }
  preprocess_small_kernel<H><<<n_windows, SMALL_THREADS, 0, stream>>>(
      colidx, rowptr, n_rows, n_windows, mode, bp, etc, etr, ht, list, counter);
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("preprocess_small launch: %s", cudaGetErrorString(err)); return (int)err; }
  // bitmap chunk: cover a square graph's column span in one chunk when it fits
  int chunk_words = (n_rows + 31) / 32;
  chunk_words = ((chunk_words + 511) / 512) * 512;
  if (chunk_words > 16384) chunk_words = 16384;
  size_t smem = (size_t)chunk_words * 8;
  cudaFuncSetAttribute(preprocess_large_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = smem > 96 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 3);
  int grid = sms * per_sm;
  if (grid > n_windows) grid = n_windows;
  preprocess_large_kernel<H><<<grid, LARGE_THREADS, smem, stream>>>(
      colidx, rowptr, n_rows, mode, chunk_words, bp, etc, ht, list, counter);
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("preprocess_large launch: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

// ranks and ceil(U/8) per 128-row super-window (dense plan, dense.cu)
int launch_preprocess_super(const int32_t *colidx, const int32_t *rowptr, int32_t n_rows, int32_t n_super,
                            int32_t *bp128, int32_t *etc128, int32_t *ht_scratch, void *ws, cudaStream_t stream) {
  return launch_preprocess_h<128>(colidx, rowptr, n_rows, n_super, HCSPMM_CLASSIFIER_ALL_CUDA, bp128, etc128,
                                  nullptr, ht_scratch, ws, stream);
}

}  // namespace hcspmm
