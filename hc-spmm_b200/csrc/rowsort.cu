// rowsort.cu -- a row-sorted copy of the CSR for the work-balanced SpMM kernel (per-graph product of preprocess).
//
// An item of the balanced kernel (csrc/spmm.cu) is a run of consecutive rows.  On a power-law graph in arbitrary
// vertex order such a run mixes rows of 1 and of 1000 entries: the lane groups of a warp that share short rows wait
// for the longest of them, and the warp-per-row phase ends ragged.  Measured on the products shape (mean row 25):
// the same graph with its rows merely GROUPED BY THE POWER OF TWO OF THEIR LENGTH (order inside a class unchanged)
// runs 15 % faster -- the whole gain of a full degree sort, and 4/5 of what relabelling rows AND columns by degree
// gives (scripts/r2/row_order_probe.py).  Row order is invisible to the caller: the kernel keeps
//     row_id[i]      original row of sorted row i          (classes in descending length, stable inside a class)
//     rowptr_s[i]    prefix of the sorted row lengths
//     colidx_s       the rows' entries, each row's in its original order (so every row sum is the same sum)
// walks (rowptr_s, colidx_s) and writes row i's result to Y[row_id[i]].  No vertex is relabelled, X is not touched.
//
// Three small kernels, no library sort: a 32-class stable counting sort (per-block class histograms with
// __match_any_sync, one scan over blocks x classes, stable scatter), a one-CTA exclusive scan of the lengths, and a
// warp-per-row copy of the entries.
#include "common.cuh"

namespace hcspmm {

namespace {
constexpr int RS_THREADS = 1024;
constexpr int RS_WARPS = RS_THREADS / 32;

// class of a row: 0 for the longest rows ... 31 for rows of 0 / 1 entries (descending length)
__device__ __forceinline__ int row_class(int len) { return len <= 1 ? 31 : __clz(len); }

// hist[block][c] = rows of class c in the block's 1024 rows
__global__ void __launch_bounds__(RS_THREADS) rowsort_hist_kernel(const int *__restrict__ rowptr, int n, int *__restrict__ hist) {
  __shared__ int s_cnt[32];
  const int r = blockIdx.x * RS_THREADS + threadIdx.x;
  if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int c = r < n ? row_class(__ldg(rowptr + r + 1) - __ldg(rowptr + r)) : -1;
  const unsigned m = __match_any_sync(0xffffffffu, c);
  if (c >= 0 && (threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&s_cnt[c], __popc(m));
  __syncthreads();
  if (threadIdx.x < 32) hist[blockIdx.x * 32 + threadIdx.x] = s_cnt[threadIdx.x];
}

// hist[block][c] -> first sorted position of the block's rows of class c (classes in order 0..31, blocks ascending)
__global__ void __launch_bounds__(32) rowsort_scan_kernel(int *__restrict__ hist, int n_blocks) {
  const int c = threadIdx.x;
  int total = 0;
  for (int b = 0; b < n_blocks; ++b) total += hist[b * 32 + c];
  int base = total;                       // exclusive prefix of the class totals over the warp
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, base, o);
    if (c >= o) base += v;
  }
  base -= total;
  for (int b = 0; b < n_blocks; ++b) {
    const int h = hist[b * 32 + c];
    hist[b * 32 + c] = base;
    base += h;
  }
}

// stable scatter: sorted position = block/class base + rows of the class in earlier warps of the block + rank in warp
__global__ void __launch_bounds__(RS_THREADS) rowsort_scatter_kernel(const int *__restrict__ rowptr, int n,
                                                                     const int *__restrict__ base, int *__restrict__ row_id,
                                                                     int *__restrict__ len_s) {
  __shared__ int s_wcnt[RS_WARPS][32];
  const int r = blockIdx.x * RS_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  s_wcnt[w][lane] = 0;
  __syncwarp();
  const int len = r < n ? __ldg(rowptr + r + 1) - __ldg(rowptr + r) : 0;
  const int c = r < n ? row_class(len) : -1;
  const unsigned m = __match_any_sync(0xffffffffu, c);
  if (c >= 0 && lane == __ffs(m) - 1) s_wcnt[w][c] = __popc(m);
  __syncthreads();
  if (c < 0) return;
  int before = 0;
  for (int ww = 0; ww < w; ++ww) before += s_wcnt[ww][c];
  const int pos = __ldg(base + blockIdx.x * 32 + c) + before + __popc(m & ((1u << lane) - 1u));
  row_id[pos] = r;
  len_s[pos] = len;
}

// rowptr_s = exclusive prefix of len_s (in place over the n + 1 entries; len_s[n] is ignored), one CTA
__global__ void __launch_bounds__(RS_THREADS) rowsort_prefix_kernel(int *__restrict__ a, int n) {
  __shared__ long long s_part[RS_THREADS];
  const int per = (n + RS_THREADS - 1) / RS_THREADS;
  const int lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < RS_THREADS; o <<= 1) {           // Hillis-Steele inclusive scan of the partial sums
    const long long v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  long long run = s_part[threadIdx.x] - sum;
  for (int i = lo; i < hi; ++i) {
    const int v = a[i];
    a[i] = (int)run;
    run += v;
  }
  if (threadIdx.x == RS_THREADS - 1) a[n] = (int)s_part[RS_THREADS - 1];
}

// colidx_s[rowptr_s[i] ...] = colidx[rowptr[row_id[i]] ...], one warp per row
__global__ void __launch_bounds__(256) rowsort_copy_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx,
                                                           const int *__restrict__ row_id, const int *__restrict__ rowptr_s,
                                                           int n, int *__restrict__ colidx_s) {
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
    const int r = __ldg(row_id + i);
    const int s = __ldg(rowptr + r), len = __ldg(rowptr + r + 1) - s, d = __ldg(rowptr_s + i);
    for (int j = lane; j < len; j += 32) colidx_s[d + j] = __ldg(colidx + s + j);
  }
}
}  // namespace

size_t row_sort_workspace_bytes(int32_t n_rows) {
  const size_t n_blocks = ((size_t)(n_rows > 0 ? n_rows : 0) + RS_THREADS - 1) / RS_THREADS;
  return sizeof(int) * 32 * (n_blocks + 1) + 256;
}

int launch_row_sort(const int32_t *rowptr, const int32_t *colidx, int32_t n, int64_t nnz, int32_t *row_id,
                    int32_t *rowptr_s, int32_t *colidx_s, void *ws, size_t ws_bytes, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (!rowptr || !row_id || !rowptr_s || (nnz > 0 && (!colidx || !colidx_s)) || !ws || ws_bytes < row_sort_workspace_bytes(n)) {
    set_error("row_sort: null pointer or workspace too small");
    return HCSPMM_E_INVALID;
  }
  const int n_blocks = (n + RS_THREADS - 1) / RS_THREADS;
  int *hist = reinterpret_cast<int *>(ws);
  rowsort_hist_kernel<<<n_blocks, RS_THREADS, 0, stream>>>(rowptr, n, hist);
  rowsort_scan_kernel<<<1, 32, 0, stream>>>(hist, n_blocks);
  rowsort_scatter_kernel<<<n_blocks, RS_THREADS, 0, stream>>>(rowptr, n, hist, row_id, rowptr_s);
  rowsort_prefix_kernel<<<1, RS_THREADS, 0, stream>>>(rowptr_s, n);
  if (nnz > 0) {
    const int grid = n / 8 + 1 < 148 * 16 ? n / 8 + 1 : 148 * 16;
    rowsort_copy_kernel<<<grid, 256, 0, stream>>>(rowptr, colidx, row_id, rowptr_s, n, colidx_s);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("row_sort: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

}  // namespace hcspmm
