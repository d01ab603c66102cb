// capi.cu -- the extern "C" surface of libhcspmm.so (declared in include/hcspmm.h).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace hcspmm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

Tuning &tuning() {
  static Tuning t = {1024, 0, 1, 1, 0, 1, 1, 2, 1, 1, 1, 0, 0, 64, 1, 0, 2048, 10000, 1, 1, 8, 72, 2048, 0};
  return t;
}

cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t stream) {
  static cudaMemPool_t pools[64] = {nullptr};
  static int keep_mb[64] = {0};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    err = cudaMemPoolCreate(&pools[dev], &props);
    if (err != cudaSuccess) { pools[dev] = nullptr; return err; }
    keep_mb[dev] = -1;
  }
  if (keep_mb[dev] != tuning().pool_keep_mb) {
    keep_mb[dev] = tuning().pool_keep_mb;
    unsigned long long keep = (unsigned long long)(keep_mb[dev] > 0 ? keep_mb[dev] : 0) << 20;
    cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
  }
  return cudaMallocFromPoolAsync(ptr, bytes ? bytes : 1, pools[dev], stream);
}

void scratch_free(void *ptr, cudaStream_t stream) {
  if (ptr) cudaFreeAsync(ptr, stream);
}

size_t preprocess_workspace_bytes(int32_t n_rows, int64_t nnz);
int launch_preprocess(const int32_t *, const int32_t *, int32_t, int64_t, int32_t, int, int32_t *,
                      int32_t *, int32_t *, int32_t *, void *, size_t, cudaStream_t);
int launch_spmm(const float *, int64_t, int32_t, const int32_t *, const int32_t *, const int32_t *,
                const int32_t *, const int32_t *, const int32_t *, int32_t, int64_t, int32_t, int,
                int, float *, int64_t, const hcspmm_aux_t *, cudaStream_t);
int launch_merge_path_splits(const int32_t *, int32_t, int64_t, int32_t, int32_t *, cudaStream_t);
int launch_f32_to_bf16(const float *, int64_t, int32_t, int32_t, void *, int64_t, cudaStream_t);
void launch_pad_rows(const float *, int64_t, int32_t, int32_t, float *, int32_t, cudaStream_t);
void launch_unpad_rows(const float *, int32_t, int32_t, int32_t, float *, int64_t, cudaStream_t);
size_t balanced_workspace_bytes(int32_t, int64_t, int32_t);
int launch_gemm_tf32(const float *, int64_t, const float *, int64_t, int32_t, int32_t, int32_t,
                     float *, int64_t, cudaStream_t);

bool umma_gemm_supported(const float *, int64_t, const float *, int64_t, int32_t, int32_t);
bool update_gemm_tma_supported(const float *, int64_t, const float *, int64_t, const float *, int64_t, int32_t, int32_t,
                               int32_t);
size_t update_gemm_scratch_floats(int32_t k, int32_t n);
int launch_update_gemm_tma(const float *, int64_t, const float *, int64_t, int32_t, int32_t, int32_t, float *, int64_t,
                           float *, int *, cudaStream_t);
int launch_umma_gemm(const float *, int64_t, const float *, int64_t, int32_t, int32_t, int32_t, float *, int64_t,
                     int *, cudaStream_t);
size_t dense_plan_workspace_bytes(int32_t n_rows, int64_t nnz);
int dense_plan_count(const int32_t *, const int32_t *, const int32_t *, int32_t, int64_t, int, void *, size_t, int32_t *,
                     cudaStream_t);
size_t dense_plan_words(int32_t n_rows, int32_t n_dense, int64_t total_cols);
size_t row_sort_workspace_bytes(int32_t n_rows);
int launch_row_sort(const int32_t *rowptr, const int32_t *colidx, int32_t n, int64_t nnz, int32_t *row_id,
                    int32_t *rowptr_s, int32_t *colidx_s, void *ws, size_t ws_bytes, cudaStream_t stream);
int dense_plan_fill(const int32_t *, const int32_t *, int32_t, int64_t, void *, int32_t, int64_t, int32_t *, size_t,
                    cudaStream_t);
bool dense_supported(const float *, const float *, int64_t, int32_t);
int launch_spmm_dense(const float *, int64_t, int32_t, int32_t, int32_t, const int32_t *, int32_t, int64_t, int, float *,
                      int64_t, float *, int *, cudaStream_t);
const int32_t *dense_plan_labels(const int32_t *, int32_t, int32_t, int64_t);
void dense_plan_arrays(const int32_t *, int32_t, int32_t, int64_t, const int **, const int **, const int **, const unsigned **);
bool dense_tma_supported(const float *, int64_t, const float *, int64_t, int32_t);
size_t dense_tma_scratch_floats(int32_t, int32_t);
int launch_spmm_dense_tma(const float *, int64_t, int32_t, int32_t, int32_t, const int *, const int *, const int *,
                          const unsigned *, int32_t, int, float *, int64_t, const float *, int64_t, int32_t, float *, int64_t,
                          float *, float *, int *, cudaStream_t);
size_t loa_workspace_bytes(int32_t n, int64_t nnz, int32_t max_degree);
int launch_loa(const int32_t *, const int32_t *, const int32_t *, const int32_t *, int32_t, int64_t, int32_t,
               int32_t *, int32_t *, int32_t *, void *, size_t, cudaStream_t);

}  // namespace hcspmm

using namespace hcspmm;

struct hcspmm_graph {
  int32_t n_rows, x_rows, n_windows;
  int64_t nnz;
  int32_t *rowptr, *colidx, *bp, *etc, *etr, *ht;
  int32_t *splits;       // per-graph products (hcspmm_aux_t): merge-path split points, label-1 window count,
  int32_t n_splits, n_tc; // and the balanced kernel's workspace (sized with the X / Y buffers)
  void *ws;
  size_t ws_bytes;
  float *x, *y;
  int32_t buf_dim;
  cudaStream_t stream;
};

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      set_error("%s: %s", #expr, cudaGetErrorString(e_));                     \
      return (int)e_;                                                         \
    }                                                                         \
  } while (0)

extern "C" {

int hcspmm_version(void) { return 100; }

const char *hcspmm_last_error(void) { return g_err; }

int hcspmm_set_tuning(const char *key, int value) {
  int *slot = nullptr;
  if (key && !strcmp(key, "long_row")) slot = &tuning().long_row;
  else if (key && !strcmp(key, "slab")) slot = &tuning().slab;
  else if (key && !strcmp(key, "vec8")) slot = &tuning().vec8;
  else if (key && !strcmp(key, "short_row")) slot = &tuning().short_row;
  else if (key && !strcmp(key, "wpc")) slot = &tuning().wpc;
  else if (key && !strcmp(key, "umma")) slot = &tuning().umma;
  else if (key && !strcmp(key, "pad_odd")) slot = &tuning().pad_odd;
  else if (key && !strcmp(key, "umma_gemm")) slot = &tuning().umma_gemm;
  else if (key && !strcmp(key, "dense_ws")) slot = &tuning().dense_ws;
  else if (key && !strcmp(key, "occupancy3")) slot = &tuning().occupancy3;
  else if (key && !strcmp(key, "balance")) slot = &tuning().balance;
  else if (key && !strcmp(key, "chunk")) slot = &tuning().chunk;
  else if (key && !strcmp(key, "warp_split")) slot = &tuning().warp_split;
  else if (key && !strcmp(key, "pull_ctas")) slot = &tuning().pull_ctas;
  else if (key && !strcmp(key, "gemm_round")) slot = &tuning().gemm_round;
  else if (key && !strcmp(key, "gemm_stages")) slot = &tuning().gemm_stages;
  else if (key && !strcmp(key, "pool_keep_mb")) slot = &tuning().pool_keep_mb;
  else if (key && !strcmp(key, "barrier_timeout_ms")) slot = &tuning().barrier_timeout_ms;
  else if (key && !strcmp(key, "dense_tma")) slot = &tuning().dense_tma;
  else if (key && !strcmp(key, "fuse_update")) slot = &tuning().fuse_update;
  else if (key && !strcmp(key, "dense_min_rowlen")) slot = &tuning().dense_min_rowlen;
  else if (key && !strcmp(key, "l2_hot_mb")) slot = &tuning().l2_hot_mb;
  else if (key && !strcmp(key, "l2_hot_min_row")) slot = &tuning().l2_hot_min_row;
  else if (key && !strcmp(key, "staged")) slot = &tuning().staged;
  if (!slot) return -1;
  int old = *slot;
  *slot = value;
  return old;
}

size_t hcspmm_preprocess_workspace_bytes(int32_t n_rows, int64_t nnz) {
  return preprocess_workspace_bytes(n_rows, nnz);
}

int hcspmm_preprocess(const int32_t *d_colidx, const int32_t *d_rowptr, int32_t n_rows, int64_t nnz,
                      int32_t n_windows, int classifier, int32_t *d_block_partition,
                      int32_t *d_edge_to_column, int32_t *d_edge_to_row, int32_t *d_hybrid_type,
                      void *d_workspace, size_t workspace_bytes, void *stream) {
  return launch_preprocess(d_colidx, d_rowptr, n_rows, nnz, n_windows, classifier,
                           d_block_partition, d_edge_to_column, d_edge_to_row, d_hybrid_type,
                           d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int hcspmm_spmm(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                const int32_t *d_colidx, const int32_t *d_block_partition,
                const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                int precision, int accumulate, float *d_y, int64_t ldy, void *stream) {
  return launch_spmm(d_x, ldx, x_rows, d_rowptr, d_colidx, d_block_partition, d_edge_to_column,
                     d_edge_to_row, d_hybrid_type, n_rows, nnz, dim, precision, accumulate, d_y,
                     ldy, nullptr, (cudaStream_t)stream);
}

size_t hcspmm_merge_path_count(int32_t n_rows, int64_t nnz, int32_t chunk) {
  if (chunk <= 0 || n_rows < 0 || nnz < 0) return 0;
  return (size_t)(((long long)n_rows + nnz + chunk - 1) / chunk) + 1;
}

int hcspmm_merge_path_splits(const int32_t *d_rowptr, int32_t n_rows, int64_t nnz, int32_t chunk, int32_t *d_splits,
                             void *stream) {
  return launch_merge_path_splits(d_rowptr, n_rows, nnz, chunk, d_splits, (cudaStream_t)stream);
}

int hcspmm_f32_to_bf16(const float *d_x, int64_t ldx, int32_t rows, int32_t dim, void *d_out, int64_t ld_out,
                       void *stream) {
  return launch_f32_to_bf16(d_x, ldx, rows, dim, d_out, ld_out, (cudaStream_t)stream);
}

size_t hcspmm_spmm_workspace_bytes(int32_t n_rows, int64_t nnz, int32_t dim) {
  return balanced_workspace_bytes(n_rows, nnz, dim);
}

static int *umma_error_flag() {
  static int *flag[16] = {nullptr};   // per device, shared by all host threads (autograd worker included)
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return nullptr;
  if (!flag[dev]) {
    if (cudaMalloc(&flag[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag[dev], 0, sizeof(int));
  }
  return flag[dev];
}

int hcspmm_gemm_tf32(const float *d_a, int64_t lda, const float *d_b, int64_t ldb, int32_t m,
                     int32_t k, int32_t n, float *d_out, int64_t ldo, void *stream) {
  if (!d_a || !d_b || !d_out) { set_error("gemm_tf32: null pointer argument"); return HCSPMM_E_INVALID; }
  if (m <= 0 || n <= 0) return 0;
  // 2 (default): TMA + tcgen05 persistent kernel; 1: register-staged tcgen05 kernel; 0: mma.sync kernel
  if (tuning().umma_gemm >= 2 && lda >= k && ldb >= n && ldo >= n && k > 0) {
    // TMA needs 16-byte aligned rows.  Operands that are not (47 classes: lda or ldo = 47) go through zero-padded
    // scratch copies -- one extra streaming pass each, after which the product runs at the HBM roofline instead of
    // on the mma.sync kernel at a third of it (only worth it for tall operands).
    cudaStream_t st = (cudaStream_t)stream;
    const bool a_ok = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(d_a) & 15) == 0;
    const bool o_ok = (ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0;
    const bool tall = (long long)m * (k + n) >= (1 << 20);
    if ((a_ok && o_ok) || tall) {
      const int32_t kp = (k + 3) / 4 * 4, np = (n + 3) / 4 * 4;
      float *wt = nullptr, *ap = nullptr, *op = nullptr;
      cudaError_t e = scratch_alloc((void **)&wt, sizeof(float) * update_gemm_scratch_floats(k, n), st);
      if (e == cudaSuccess && !a_ok) e = scratch_alloc((void **)&ap, sizeof(float) * (size_t)m * kp, st);
      if (e == cudaSuccess && !o_ok) e = scratch_alloc((void **)&op, sizeof(float) * (size_t)m * np, st);
      int rc = 0;
      if (e != cudaSuccess) { set_error("gemm_tf32: scratch: %s", cudaGetErrorString(e)); rc = (int)e; }
      if (rc == 0 && !a_ok) launch_pad_rows(d_a, lda, m, k, ap, kp, st);
      if (rc == 0 && update_gemm_tma_supported(a_ok ? d_a : ap, a_ok ? lda : kp, d_b, ldb, o_ok ? d_out : op, o_ok ? ldo : np, m, k, n)) {
        rc = launch_update_gemm_tma(a_ok ? d_a : ap, a_ok ? lda : kp, d_b, ldb, m, k, n, o_ok ? d_out : op, o_ok ? ldo : np, wt,
                                    umma_error_flag(), st);
        if (rc == 0 && !o_ok) launch_unpad_rows(op, np, m, n, d_out, ldo, st);
      } else if (rc == 0) {
        rc = launch_gemm_tf32(d_a, lda, d_b, ldb, m, k, n, d_out, ldo, st);
      }
      if (wt) scratch_free(wt, st);
      if (ap) scratch_free(ap, st);
      if (op) scratch_free(op, st);
      return rc;
    }
  }
  if (tuning().umma_gemm == 1 && lda >= k && ldb >= n && ldo >= n && umma_gemm_supported(d_a, lda, d_b, ldb, k, n))
    return launch_umma_gemm(d_a, lda, d_b, ldb, m, k, n, d_out, ldo, umma_error_flag(), (cudaStream_t)stream);
  return launch_gemm_tf32(d_a, lda, d_b, ldb, m, k, n, d_out, ldo, (cudaStream_t)stream);
}

size_t hcspmm_row_sort_workspace_bytes(int32_t n_rows) { return row_sort_workspace_bytes(n_rows); }

int hcspmm_row_sort(const int32_t *d_rowptr, const int32_t *d_colidx, int32_t n_rows, int64_t nnz, int32_t *d_row_id,
                    int32_t *d_sorted_rowptr, int32_t *d_sorted_colidx, void *d_workspace, size_t workspace_bytes,
                    void *stream) {
  if (n_rows < 0 || nnz < 0) { set_error("row_sort: negative size"); return HCSPMM_E_INVALID; }
  return launch_row_sort(d_rowptr, d_colidx, n_rows, nnz, d_row_id, d_sorted_rowptr, d_sorted_colidx, d_workspace,
                         workspace_bytes, (cudaStream_t)stream);
}

size_t hcspmm_dense_plan_workspace_bytes(int32_t n_rows, int64_t nnz) { return dense_plan_workspace_bytes(n_rows, nnz); }

int hcspmm_dense_plan_count(const int32_t *d_colidx, const int32_t *d_rowptr, const int32_t *d_hybrid_type,
                            int32_t n_rows, int64_t nnz, int min_reuse_x2, void *d_workspace, size_t workspace_bytes,
                            int32_t *h_counts, void *stream) {
  if (!h_counts) { set_error("dense_plan_count: null h_counts"); return HCSPMM_E_INVALID; }
  return dense_plan_count(d_colidx, d_rowptr, d_hybrid_type, n_rows, nnz, min_reuse_x2, d_workspace, workspace_bytes,
                          h_counts, (cudaStream_t)stream);
}

size_t hcspmm_dense_plan_words(int32_t n_rows, int32_t n_dense, int64_t total_cols) {
  return dense_plan_words(n_rows, n_dense, total_cols);
}

int hcspmm_dense_plan_fill(const int32_t *d_colidx, const int32_t *d_edge_to_row, int32_t n_rows, int64_t nnz,
                           void *d_workspace, int32_t n_dense, int64_t total_cols, int32_t *d_plan, size_t plan_words,
                           void *stream) {
  if (!d_colidx || !d_edge_to_row || !d_workspace || !d_plan) { set_error("dense_plan_fill: null pointer argument"); return HCSPMM_E_INVALID; }
  return dense_plan_fill(d_colidx, d_edge_to_row, n_rows, nnz, d_workspace, n_dense, total_cols, d_plan, plan_words,
                         (cudaStream_t)stream);
}

int hcspmm_spmm_aux(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                    const int32_t *d_colidx, const int32_t *d_block_partition,
                    const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                    const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                    int precision, int accumulate, float *d_y, int64_t ldy, const hcspmm_aux_t *aux, void *stream) {
  const int32_t *d_plan = aux ? aux->d_plan : nullptr;
  const int32_t n_dense = aux ? aux->n_dense : 0;
  const int64_t total_cols = aux ? aux->total_cols : 0;
  // the tcgen05 kernel holds a [128 x <=256] accumulator in TMEM: wider operands go in column blocks of 256
  const int32_t dblock = dim > 256 ? 256 : dim;
  if (aux && aux->d_colidx_segments && d_plan && n_dense > 0) {
    set_error("spmm: segment-tagged column ids cannot be combined with a dense super-window plan");
    return HCSPMM_E_INVALID;
  }
  const bool dense = d_plan && n_dense > 0 && tuning().umma && precision == HCSPMM_PRECISION_TF32 && d_x && d_y &&
                     (ldx & 3) == 0 && (dim % 16) == 0 && dense_supported(d_x, d_y, ldy, dblock);
  if (!dense)
    return launch_spmm(d_x, ldx, x_rows, d_rowptr, d_colidx, d_block_partition, d_edge_to_column, d_edge_to_row,
                       d_hybrid_type, n_rows, nnz, dim, precision, accumulate, d_y, ldy, aux, (cudaStream_t)stream);
  int rc = 0;
  // "dense_tma": 0 / 1 the kernels of dense.cu (1, default: the fused entry point still uses dense_tma.cu);
  //              2 dense_tma.cu with cp.async gathers; 3 dense_tma.cu with TMA gather4 (no copy of X)
  const int mode = tuning().dense_tma >= 2 ? tuning().dense_tma - 1 : 0;
  float *xr = nullptr;
  if (mode != 2 || !dense_tma_supported(d_x, ldx, d_y, ldy, dblock)) {
    cudaError_t err = scratch_alloc((void **)&xr, sizeof(float) * (size_t)x_rows * dblock, (cudaStream_t)stream);
    if (err != cudaSuccess) { set_error("spmm_plan: scratch: %s", cudaGetErrorString(err)); return (int)err; }
  }
  if (mode >= 1 && dense_tma_supported(d_x, ldx, d_y, ldy, dblock)) {
    const int *sw_ids, *sw_off, *cols;
    const unsigned *masks;
    dense_plan_arrays(d_plan, n_rows, n_dense, total_cols, &sw_ids, &sw_off, &cols, &masks);
    for (int32_t c0 = 0; c0 < dim && rc == 0; c0 += dblock) {
      const int32_t w = dim - c0 < dblock ? dim - c0 : dblock;
      rc = launch_spmm_dense_tma(d_x + c0, ldx, x_rows, n_rows, w, sw_ids, sw_off, cols, masks, n_dense, accumulate,
                                 d_y + c0, ldy, nullptr, 0, 0, nullptr, 0, nullptr, xr, umma_error_flag(),
                                 (cudaStream_t)stream);
    }
  } else {
    for (int32_t c0 = 0; c0 < dim && rc == 0; c0 += dblock) {
      const int32_t w = dim - c0 < dblock ? dim - c0 : dblock;
      rc = launch_spmm_dense(d_x + c0, ldx, x_rows, n_rows, w, d_plan, n_dense, total_cols, accumulate, d_y + c0, ldy, xr,
                             umma_error_flag(), (cudaStream_t)stream);
    }
  }
  if (xr) scratch_free(xr, (cudaStream_t)stream);
  if (rc == 0 && aux->plan_full) {
    // every row with stored entries was computed above: nothing is left for the per-window / balanced kernels
    // (their launches cost 0.14 ms per aggregation on the proteins shape just to skip every window); rows of
    // super-windows without entries are zero
    if ((long long)n_dense * 128 < n_rows && !accumulate) {
      cudaError_t e = cudaMemset2DAsync(d_y, sizeof(float) * ldy, 0, sizeof(float) * dim, n_rows, (cudaStream_t)stream);
      if (e != cudaSuccess) { set_error("spmm_plan: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
  } else if (rc == 0) {
    hcspmm_aux_t rest = *aux;
    rest.n_tc_windows = -1;   // the plan's labels differ from the caller's count
    rc = launch_spmm(d_x, ldx, x_rows, d_rowptr, d_colidx, d_block_partition, d_edge_to_column, d_edge_to_row,
                     dense_plan_labels(d_plan, n_rows, n_dense, total_cols), n_rows, nnz, dim, precision, accumulate,
                     d_y, ldy, &rest, (cudaStream_t)stream);
  }
  return rc;
}

int hcspmm_spmm_plan(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                     const int32_t *d_colidx, const int32_t *d_block_partition,
                     const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                     const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                     int precision, int accumulate, float *d_y, int64_t ldy, const int32_t *d_plan,
                     int32_t n_dense, int64_t total_cols, void *stream) {
  hcspmm_aux_t aux;
  memset(&aux, 0, sizeof(aux));
  aux.n_tc_windows = -1;
  aux.d_plan = d_plan; aux.n_dense = n_dense; aux.total_cols = total_cols;
  return hcspmm_spmm_aux(d_x, ldx, x_rows, d_rowptr, d_colidx, d_block_partition, d_edge_to_column, d_edge_to_row,
                         d_hybrid_type, n_rows, nnz, dim, precision, accumulate, d_y, ldy, &aux, stream);
}

int hcspmm_debug_umma_error(void) {
  int *f = umma_error_flag();
  if (!f) return -1;
  int v = 0;
  if (cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (v) cudaMemset(f, 0, sizeof(int));
  return v;
}

int hcspmm_spmm_gemm(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                     const int32_t *d_colidx, const int32_t *d_block_partition,
                     const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                     const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                     int precision, const float *d_w, int64_t ldw, int32_t hidden, float *d_out,
                     int64_t ldo, float *d_z, int64_t ldz, void *stream) {
  if (!d_z || !d_out || !d_w) {
    set_error("spmm_gemm: null pointer argument");
    return HCSPMM_E_INVALID;
  }
  int rc = launch_spmm(d_x, ldx, x_rows, d_rowptr, d_colidx, d_block_partition, d_edge_to_column,
                       d_edge_to_row, d_hybrid_type, n_rows, nnz, dim, precision, 0, d_z, ldz,
                       nullptr, (cudaStream_t)stream);
  if (rc) return rc;
  return hcspmm_gemm_tf32(d_z, ldz, d_w, ldw, n_rows, dim, hidden, d_out, ldo, stream);
}

int hcspmm_spmm_gemm_aux(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                         const int32_t *d_colidx, const int32_t *d_block_partition,
                         const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                         const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                         int precision, const float *d_w, int64_t ldw, int32_t hidden, float *d_out,
                         int64_t ldo, float *d_z, int64_t ldz, const hcspmm_aux_t *aux, void *stream) {
  if (!d_z || !d_out || !d_w) {
    set_error("spmm_gemm: null pointer argument");
    return HCSPMM_E_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool fuse = aux && aux->d_plan && aux->n_dense > 0 && aux->plan_full && tuning().umma && tuning().dense_tma &&
                    tuning().fuse_update && precision == HCSPMM_PRECISION_TF32 && d_x && hidden > 0 && hidden <= 256 &&
                    dim <= 256 && ldw >= hidden && ldo >= hidden && dense_tma_supported(d_x, ldx, d_z, ldz, dim);
  if (fuse) {
    // ONE kernel: the aggregate of each super-window is multiplied by W before it leaves tensor memory
    const int *sw_ids, *sw_off, *cols;
    const unsigned *masks;
    dense_plan_arrays(aux->d_plan, n_rows, aux->n_dense, aux->total_cols, &sw_ids, &sw_off, &cols, &masks);
    if ((long long)aux->n_dense * 128 < n_rows) {   // super-windows without entries: their rows of Z and out are zero
      CUDA_TRY(cudaMemset2DAsync(d_z, sizeof(float) * ldz, 0, sizeof(float) * dim, n_rows, st));
      CUDA_TRY(cudaMemset2DAsync(d_out, sizeof(float) * ldo, 0, sizeof(float) * hidden, n_rows, st));
    }
    float *wt = nullptr, *xr = nullptr;
    cudaError_t e = scratch_alloc((void **)&wt, sizeof(float) * dense_tma_scratch_floats(dim, hidden), st);
    if (e == cudaSuccess && tuning().dense_tma != 3) e = scratch_alloc((void **)&xr, sizeof(float) * (size_t)x_rows * dim, st);
    if (e != cudaSuccess) { if (wt) scratch_free(wt, st); set_error("spmm_gemm: scratch: %s", cudaGetErrorString(e)); return (int)e; }
    const int rc = launch_spmm_dense_tma(d_x, ldx, x_rows, n_rows, dim, sw_ids, sw_off, cols, masks, aux->n_dense, 0, d_z,
                                         ldz, d_w, ldw, hidden, d_out, ldo, wt, xr, umma_error_flag(), st);
    scratch_free(wt, st);
    if (xr) scratch_free(xr, st);
    return rc;
  }
  int rc = hcspmm_spmm_aux(d_x, ldx, x_rows, d_rowptr, d_colidx, d_block_partition, d_edge_to_column, d_edge_to_row,
                           d_hybrid_type, n_rows, nnz, dim, precision, 0, d_z, ldz, aux, stream);
  if (rc) return rc;
  return hcspmm_gemm_tf32(d_z, ldz, d_w, ldw, n_rows, dim, hidden, d_out, ldo, stream);
}

size_t hcspmm_loa_workspace_bytes(int32_t n, int64_t nnz, int32_t max_degree) {
  return loa_workspace_bytes(n, nnz, max_degree);
}

int hcspmm_loa_reorder(const int32_t *d_rowptr, const int32_t *d_colidx, const int32_t *d_rowptr_in,
                       const int32_t *d_colidx_in, int32_t n, int64_t nnz, int32_t max_degree,
                       int32_t *d_perm, int32_t *d_block_start, int32_t *d_counts, void *d_workspace,
                       size_t workspace_bytes, void *stream) {
  return launch_loa(d_rowptr, d_colidx, d_rowptr_in, d_colidx_in, n, nnz, max_degree, d_perm,
                    d_block_start, d_counts, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

void hcspmm_graph_destroy(hcspmm_graph_t *g) {
  if (!g) return;
  cudaFree(g->rowptr); cudaFree(g->colidx); cudaFree(g->bp); cudaFree(g->etc);
  cudaFree(g->etr); cudaFree(g->ht); cudaFree(g->x); cudaFree(g->y); cudaFree(g->splits); cudaFree(g->ws);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
}

int hcspmm_graph_create(const int32_t *h_rowptr, const int32_t *h_colidx, int32_t n_rows,
                        int64_t nnz, int32_t x_rows, int classifier, hcspmm_graph_t **out) {
  if (!out || !h_rowptr || n_rows < 0 || nnz < 0 || (nnz > 0 && !h_colidx)) {
    set_error("graph_create: bad argument");
    return HCSPMM_E_INVALID;
  }
  hcspmm_graph *g = new hcspmm_graph();
  memset(g, 0, sizeof(*g));
  g->n_rows = n_rows; g->x_rows = x_rows; g->nnz = nnz;
  g->n_windows = (n_rows + HCSPMM_BLK_H - 1) / HCSPMM_BLK_H;
  void *ws = nullptr;
  size_t ws_bytes = preprocess_workspace_bytes(n_rows, nnz);
  size_t e = (size_t)(nnz > 0 ? nnz : 1), w = (size_t)(g->n_windows > 0 ? g->n_windows : 1);
  int rc = 0;
#define G_TRY(expr)                                                           \
  do {                                                                        \
    cudaError_t e_ = (expr);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      set_error("%s: %s", #expr, cudaGetErrorString(e_));                     \
      rc = (int)e_;                                                           \
      goto fail;                                                              \
    }                                                                         \
  } while (0)
  G_TRY(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
  G_TRY(cudaMalloc(&g->rowptr, sizeof(int32_t) * ((size_t)n_rows + 1)));
  G_TRY(cudaMalloc(&g->colidx, sizeof(int32_t) * e));
  G_TRY(cudaMalloc(&g->bp, sizeof(int32_t) * w));
  G_TRY(cudaMalloc(&g->etc, sizeof(int32_t) * e));
  G_TRY(cudaMalloc(&g->etr, sizeof(int32_t) * e));
  G_TRY(cudaMalloc(&g->ht, sizeof(int32_t) * w));
  G_TRY(cudaMalloc(&ws, ws_bytes));
  G_TRY(cudaMemcpyAsync(g->rowptr, h_rowptr, sizeof(int32_t) * ((size_t)n_rows + 1),
                        cudaMemcpyHostToDevice, g->stream));
  if (nnz > 0)
    G_TRY(cudaMemcpyAsync(g->colidx, h_colidx, sizeof(int32_t) * (size_t)nnz,
                          cudaMemcpyHostToDevice, g->stream));
  rc = launch_preprocess(g->colidx, g->rowptr, n_rows, nnz, g->n_windows, classifier, g->bp,
                         g->etc, g->etr, g->ht, ws, ws_bytes, g->stream);
  if (rc) goto fail;
  {
    const size_t cnt = hcspmm_merge_path_count(n_rows, nnz, HCSPMM_SPLITS_CHUNK);
    G_TRY(cudaMalloc(&g->splits, sizeof(int32_t) * cnt));
    rc = launch_merge_path_splits(g->rowptr, n_rows, nnz, HCSPMM_SPLITS_CHUNK, g->splits, g->stream);
    if (rc) goto fail;
    g->n_splits = (int32_t)cnt - 1;
    int32_t *h_ht = new int32_t[w];
    cudaError_t e_ = cudaMemcpyAsync(h_ht, g->ht, sizeof(int32_t) * (size_t)g->n_windows, cudaMemcpyDeviceToHost, g->stream);
    if (e_ == cudaSuccess) e_ = cudaStreamSynchronize(g->stream);
    g->n_tc = 0;
    for (int32_t i = 0; i < g->n_windows; ++i) g->n_tc += h_ht[i] == 1;
    delete[] h_ht;
    if (e_ != cudaSuccess) { set_error("graph_create: %s", cudaGetErrorString(e_)); rc = (int)e_; goto fail; }
  }
  cudaFree(ws);
  *out = g;
  return 0;
fail:
  cudaFree(ws);
  hcspmm_graph_destroy(g);
  return rc;
#undef G_TRY
}

int hcspmm_graph_spmm_host(hcspmm_graph_t *g, const float *h_x, int32_t dim, int precision,
                           float *h_y) {
  if (!g || !h_x || !h_y || dim <= 0) {
    set_error("graph_spmm_host: bad argument");
    return HCSPMM_E_INVALID;
  }
  if (dim != g->buf_dim) {
    cudaFree(g->x); cudaFree(g->y); cudaFree(g->ws);
    g->x = g->y = nullptr; g->ws = nullptr; g->buf_dim = 0;
    g->ws_bytes = balanced_workspace_bytes(g->n_rows, g->nnz, dim);
    CUDA_TRY(cudaMalloc(&g->ws, g->ws_bytes));
    CUDA_TRY(cudaMalloc(&g->x, sizeof(float) * (size_t)g->x_rows * dim));
    CUDA_TRY(cudaMalloc(&g->y, sizeof(float) * (size_t)(g->n_rows > 0 ? g->n_rows : 1) * dim));
    g->buf_dim = dim;
  }
  CUDA_TRY(cudaMemcpyAsync(g->x, h_x, sizeof(float) * (size_t)g->x_rows * dim,
                           cudaMemcpyHostToDevice, g->stream));
  hcspmm_aux_t aux;
  memset(&aux, 0, sizeof(aux));
  aux.d_splits = g->splits; aux.splits_chunk = HCSPMM_SPLITS_CHUNK; aux.n_splits = g->n_splits;
  aux.n_tc_windows = g->n_tc; aux.d_workspace = g->ws; aux.workspace_bytes = g->ws_bytes;
  int rc = launch_spmm(g->x, dim, g->x_rows, g->rowptr, g->colidx, g->bp, g->etc, g->etr, g->ht,
                       g->n_rows, g->nnz, dim, precision, 0, g->y, dim, &aux, g->stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_y, g->y, sizeof(float) * (size_t)g->n_rows * dim,
                           cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  return 0;
}

int hcspmm_graph_get_preprocess(hcspmm_graph_t *g, int32_t *h_bp, int32_t *h_etc, int32_t *h_etr,
                                int32_t *h_ht) {
  if (!g) { set_error("graph_get_preprocess: null graph"); return HCSPMM_E_INVALID; }
  if (h_bp) CUDA_TRY(cudaMemcpy(h_bp, g->bp, sizeof(int32_t) * (size_t)g->n_windows, cudaMemcpyDeviceToHost));
  if (h_ht) CUDA_TRY(cudaMemcpy(h_ht, g->ht, sizeof(int32_t) * (size_t)g->n_windows, cudaMemcpyDeviceToHost));
  if (h_etc) CUDA_TRY(cudaMemcpy(h_etc, g->etc, sizeof(int32_t) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
  if (h_etr) CUDA_TRY(cudaMemcpy(h_etr, g->etr, sizeof(int32_t) * (size_t)g->nnz, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
