// hotcols.cu -- column hotness classes for the L2 residency hints of the balanced SpMM kernel (spmm.cu, GatherHint).
//
// A per-graph product, computed once: how often every column of A is referenced (= how often its X row is
// gathered per aggregation), the columns ranked by that count, and the column ids of the CSR re-emitted with the
// rank's class in their top three bits:
//     class 7: rank < 16 384      6: < 32 768      5: < 65 536      4: < 131 072
//           3: < 262 144          2: < 524 288     1: < 1 048 576   0: the rest
// so that a kernel that knows how many rows of the current width fit its L2 budget keeps exactly the hottest ones
// resident (class >= cls_min -> evict_last, else evict_first) without any lookup per stored entry.
// Column ids must be < 2^29.  Ties in the count are broken by the radix sort's stability (ascending column id).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace hcspmm {

__global__ void count_columns_kernel(const int *__restrict__ colidx, long long nnz, int n_cols, int *__restrict__ cnt) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
    const int c = __ldg(colidx + e);
    if ((unsigned)c < (unsigned)n_cols) atomicAdd(cnt + c, 1);
  }
}
__global__ void iota_kernel(int *v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}
__global__ void rank_class_kernel(const int *__restrict__ order, int n_cols, unsigned char *__restrict__ cls) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cols) return;
  int c = 7;
  for (int k = 16384; c > 0 && i >= k; k <<= 1) --c;
  cls[order[i]] = (unsigned char)c;
}
__global__ void tag_columns_kernel(const int *__restrict__ colidx, long long nnz, int n_cols,
                                   const unsigned char *__restrict__ cls, int *__restrict__ tagged) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
    const int c = __ldg(colidx + e);
    const unsigned k = (unsigned)c < (unsigned)n_cols ? (unsigned)cls[c] : 0u;
    tagged[e] = (int)(((unsigned)c & 0x1fffffffu) | (k << 29));
  }
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sort_temp_bytes(int n_cols) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                            (int *)nullptr, n_cols);
  return bytes;
}

}  // namespace hcspmm

using namespace hcspmm;

extern "C" {

size_t hcspmm_tag_columns_workspace_bytes(int32_t n_cols, int64_t nnz) {
  (void)nnz;
  if (n_cols <= 0) return 256;
  const size_t n = (size_t)n_cols;
  return 4 * al256(n * 4) + al256(n) + al256(sort_temp_bytes(n_cols)) + 256;
}

int hcspmm_tag_columns(const int32_t *d_colidx, int64_t nnz, int32_t n_cols, int32_t *d_tagged, void *d_workspace,
                       size_t workspace_bytes, void *stream) {
  if (nnz < 0 || n_cols < 0) { set_error("tag_columns: negative size"); return HCSPMM_E_INVALID; }
  if (nnz == 0) return 0;
  if (!d_colidx || !d_tagged || !d_workspace) { set_error("tag_columns: null pointer argument"); return HCSPMM_E_INVALID; }
  if (n_cols > (1 << 29)) { set_error("tag_columns: column ids must be below 2^29"); return HCSPMM_E_UNSUPPORTED; }
  if (workspace_bytes < hcspmm_tag_columns_workspace_bytes(n_cols, nnz)) { set_error("tag_columns: workspace too small"); return HCSPMM_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)n_cols;
  char *w = reinterpret_cast<char *>(d_workspace);
  int *cnt = reinterpret_cast<int *>(w); w += al256(n * 4);
  int *cnt_sorted = reinterpret_cast<int *>(w); w += al256(n * 4);
  int *ids = reinterpret_cast<int *>(w); w += al256(n * 4);
  int *order = reinterpret_cast<int *>(w); w += al256(n * 4);
  unsigned char *cls = reinterpret_cast<unsigned char *>(w); w += al256(n);
  size_t temp = sort_temp_bytes(n_cols);
  cudaError_t err = cudaMemsetAsync(cnt, 0, n * 4, st);
  if (err == cudaSuccess) {
    count_columns_kernel<<<1184, 256, 0, st>>>(d_colidx, nnz, n_cols, cnt);
    iota_kernel<<<(n_cols + 255) / 256, 256, 0, st>>>(ids, n_cols);
    err = cub::DeviceRadixSort::SortPairsDescending(w, temp, cnt, cnt_sorted, ids, order, n_cols, 0, 32, st);
  }
  if (err == cudaSuccess) {
    rank_class_kernel<<<(n_cols + 255) / 256, 256, 0, st>>>(order, n_cols, cls);
    tag_columns_kernel<<<1184, 256, 0, st>>>(d_colidx, nnz, n_cols, cls, d_tagged);
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) { set_error("tag_columns: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

}  // extern "C"
