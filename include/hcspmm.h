/*
 * hcspmm.h -- C ABI of the B200-native HC-SpMM hot path (libhcspmm.so).
 *
 * Drop-in boundary for the reference's operator layer: every entry point below is
 * what a binding of the reference's `HCSPMM` extension would call in place of the
 * C++ launchers in /root/reference/hybrid_kernel/hybrid_all_kernel.cu.  Plain
 * pointers and sizes only -- no torch types.  All `d_*` pointers are DEVICE
 * pointers on the current CUDA device; `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream, which is what the reference launches on,
 * hybrid_all_kernel.cu:438).
 *
 * Conventions
 *   - A is a binary CSR adjacency (no values array, like the reference):
 *     rowptr int32[n_rows+1], colidx int32[nnz] (dataset.py:93-103).
 *   - The operator is rectangular: n_rows output rows, x_rows rows of X; a column
 *     id outside [0, x_rows) contributes 0 (row-partitioned multi-GPU shards keep
 *     GLOBAL column ids).
 *   - Row windows are 16 rows (BLK_H), condensed column blocks are 8 wide
 *     (BLK_W): hybrid_kernel/config.h:4-5.
 *   - Return value: 0 on success; a positive value is a cudaError_t; a negative
 *     value is an argument error (HCSPMM_E_*).  hcspmm_last_error() returns a
 *     thread-local description of the last failure.
 *   - No CPU fallback exists.  Without a usable CUDA device every compute entry
 *     point fails with a cudaError_t.
 */
#ifndef HCSPMM_H_
#define HCSPMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCSPMM_BLK_H 16
#define HCSPMM_BLK_W 8

#define HCSPMM_E_INVALID  (-1) /* bad size / null pointer            */
#define HCSPMM_E_ALIGN    (-2) /* pointer or leading dim misaligned  */
#define HCSPMM_E_WORKSPACE (-3) /* workspace too small               */
#define HCSPMM_E_UNSUPPORTED (-4)

/* Core selector (hybrid_type) variants.  0/1 are bit-exact restatements of the
 * reference, hybrid_all_kernel.cu:262 (shipped) and :261 (intended, commented
 * out in the reference).  2 is the B200 re-fit: it labels tensor-core
 * CANDIDATES (value 3 in hybrid_type), never 1.                                */
#define HCSPMM_CLASSIFIER_SHIPPED   0
#define HCSPMM_CLASSIFIER_INTENDED  1
#define HCSPMM_CLASSIFIER_B200      2
#define HCSPMM_CLASSIFIER_ALL_CUDA  3
#define HCSPMM_CLASSIFIER_ALL_TC    4
/* 5: the reference's logistic form (:261) with coefficients RE-FITTED ON B200 by the paper's recipe (16-row synthetic
 *    windows, CUDA-core path vs per-window mma.sync path, dim 128): label 1 iff
 *    -0.02312523 (U-1) - 9.74306426 density + 4.93743285 <= 0.  benchmarks/selector_fit.py, profiles/r2_selector_fit.json */
#define HCSPMM_CLASSIFIER_B200_WINDOW 5

/* Arithmetic of windows labelled "tensor core" (hybrid_type != 0).  CUDA-core
 * windows are always an exact FP32 sum, as in the reference (:982-990).
 *   TF32   : X rounded cvt.rna to TF32, FP32 accumulate -- the reference (:1102-1111)
 *   TF32X2 : X split hi+lo into two TF32 MMAs (FP32-accurate result on tensor cores)
 *   FP32   : ignore labels, every window on the CUDA-core path                      */
#define HCSPMM_PRECISION_TF32    0
#define HCSPMM_PRECISION_TF32X2  1
#define HCSPMM_PRECISION_FP32    2
/*   BF16   : X is copied to bfloat16 (round to nearest even) for the gather -- half the bytes of the
 *            dominant stream -- and accumulated in FP32 on the CUDA-core path for every window
 *            (north star: 1e-2 against FP32 torch.sparse).  Needs dim % 8 == 0, else computed in FP32. */
#define HCSPMM_PRECISION_BF16    3
/*   BF16_STORED : as BF16, but d_x ALREADY holds bfloat16 rows (ldx counted in bfloat16 elements; dim and ldx
 *            multiples of 8, 16-byte aligned): the operand of a multi-GPU aggregation whose halo rows travelled
 *            as bfloat16 (hcspmm_f32_to_bf16 writes a rank's own rows into it).  No conversion pass.           */
#define HCSPMM_PRECISION_BF16_STORED 4

int hcspmm_version(void);
const char *hcspmm_last_error(void);

/* Runtime tuning knobs (process-wide; for experiments and tests).
 *   "long_row"   rows with >= this many non-zeros are split over all warps of a CTA
 *   "slab"       feature-slab width in floats (0 = whole row); slabs are scheduled
 *                slab-major so that one slab of X stays L2-resident
 *   "vec8"       1 (default): 256-bit gathers when rows are 32-byte aligned; 0: 128-bit
 *   "short_row"  rows with fewer than short_row * G entries (G = rows a warp can advance at once,
 *                32 / lanes-per-row) are processed one lane group per row; 0 disables
 *   "wpc"        16-row windows per CTA (1..8); 0 = chosen from nnz / windows
 *   "pad_odd"    1 (default): large operands whose width is not a multiple of 4 (or whose rows
 *                are unaligned) run through zero-padded aligned copies; 0: scalar kernel
 *   "umma"       1: tcgen05 / TMEM dense super-window kernel (hcspmm_spmm_plan); 0: per-window paths
 *   "umma_gemm"  Update GEMM kernel: 2 = TMA + tcgen05 persistent warp-specialised kernel, 1 = register-staged
 *                tcgen05 kernel, 0 = mma.sync kernel (also the fallback for unaligned operands)
 *   "dense_min_rowlen" a super-window joins the dense plan only if it holds >= this many stored entries per row on
 *                average (default 8: the boundary the B200 re-fit found, profiles/r2_selector_fit.json); the
 *                min_reuse_x2 argument of hcspmm_dense_plan_count applies as well (0 = force: no threshold at all)
 *   "l2_hot_mb"  megabytes of gathered X rows the balanced kernel keeps L2-resident when the caller passes tagged
 *                column ids (default 72 of the 126 MB L2; 0 = hints off)
 *   "l2_hot_min_row" the hints apply to gathers of at least this many bytes per row (default 2048: measured to lose
 *                below -- LRU already keeps the hub rows -- and to gain 7 % at the Reddit shape, dim 512)
 *   "staged"     1: low-degree graphs (mean row < 64) with rows of 260..512 bytes and no tensor-core window gather
 *                through shared-memory staging (spmm_staged_kernel: cp.async ring of 20 rows per warp, one lane per
 *                16 bytes of a row, no register held by a row in flight) instead of the register ring
 *   "dense_tma"  which kernel multiplies dense super-windows.  csrc/dense_tma.cu is the five-role kernel (TMA for
 *                the plan's index chunks and W^T, dedicated epilogue warps, optional FUSED Update); csrc/dense.cu holds
 *                the earlier producer/issuer kernels.  1 (default): dense.cu for plain aggregation (measured fastest,
 *                proteins shape dim 256: 0.64 ms), dense_tma.cu with cp.async gathers behind the fused entry point
 *                (0.62 ms for Z and out together, against 0.70 ms unfused); 2: dense_tma.cu with cp.async gathers for
 *                both; 3: dense_tma.cu gathering X rows with TMA gather4 from a TF32-typed tensor map (no rounded copy
 *                of X, but the TMA unit serves one 512-byte gather4 per ~46 cycles: 1.28 ms); 0: dense.cu only, the
 *                fused entry point runs aggregation and Update GEMM back to back
 *   "fuse_update" 1 (default): hcspmm_spmm_gemm_aux fuses the Update product when the plan covers the graph
 *   "gemm_round" TMA Update GEMM: 1 (default) rounder warps apply cvt.rna.tf32 to the landed Z boxes (the
 *                reference's rounding); 0 = the tensor map's TF32 element type converts on load
 *   "gemm_stages" cap on the TMA Update GEMM's shared-memory ring depth (0 = as many stages as fit)
 *   "barrier_timeout_ms" how long hcspmm_peer_barrier waits for a peer (default 10000) before it sets *d_err
 *   "pool_keep_mb" megabytes of freed scratch the library's private stream-ordered pool keeps mapped
 *   "balance"    CUDA-core windows on the merge-path balanced kernel (equal rows + stored entries per
 *                CTA, hub rows cut into pieces that are summed in a fixed order): 1 (default) when the
 *                mean row holds >= 8 entries, 2 always, 0 never (one CTA per "wpc" 16-row windows)
 *   "chunk"      rows + stored entries per item of the balanced kernel (0 = about 4 MB of gathered rows)
 *   "pull_ctas"  grid cap of hcspmm_halo_pull (0 = 1184)
 *   "occupancy3" low-degree graphs (mean row < 64 entries) use a build of the SpMM kernels that fits 3 (value 1,
 *                default) or 4 (value 2) CTAs per SM instead of 2; 0 = always the 2-CTA build
 *   "dense_ws"   1 (default): warp-specialised tcgen05 dense kernel; 0: the single-role variant
 *   "warp_split" items of the balanced kernel whose mean row length is >= this (default 64) give every warp
 *                an equal run of entries; 0 = warp-per-row / CTA-per-long-row phases everywhere
 * Returns the previous value, or -1 for an unknown key.                          */
int hcspmm_set_tuning(const char *key, int value);

/* ---- A1-A5: preprocessing -------------------------------------------------------
 * Replaces preprocess(), hybrid_all_kernel.cu:339-408 (pybind: hybrid_all.cpp:501):
 * fill_edgeToRow (:314-326), fill_segment + thrust::sort (:289-301, :386-399) and
 * generate_edgetocolumn (:242-269), as one window-parallel pass (no global sort).
 * Outputs are bit-exact with the reference for every non-empty window; empty
 * windows are defined as 0 (the reference leaves them uninitialised, :252-253,
 * :356-366).  d_workspace: hcspmm_preprocess_workspace_bytes() bytes of scratch.   */
size_t hcspmm_preprocess_workspace_bytes(int32_t n_rows, int64_t nnz);
int hcspmm_preprocess(const int32_t *d_colidx, const int32_t *d_rowptr, int32_t n_rows,
                      int64_t nnz, int32_t n_windows, int classifier,
                      int32_t *d_block_partition, int32_t *d_edge_to_column,
                      int32_t *d_edge_to_row, int32_t *d_hybrid_type, void *d_workspace,
                      size_t workspace_bytes, void *stream);

/* ---- A6/A7: Y = A * X ------------------------------------------------------------
 * Replaces spmm_forward_plus / _more / _fixed32 / _fixed64 and their kernels
 * (hybrid_all_kernel.cu:410-595, 919-1637; pybind hybrid_all.cpp:194-308) for ANY
 * dim (the reference needs dim == 32 / 64 / <= 48 depending on the entry point).
 * Each 16-row window takes the CUDA-core or the tensor-core path by d_hybrid_type[w]
 * (0 = CUDA cores, 1 = tensor cores as in the reference; 3 = tensor-core candidate of the
 * B200 selector: CUDA cores here, tcgen05 under hcspmm_spmm_plan).  CUDA-core windows run
 * on the work-balanced kernel (merge_path_splits + spmm_balanced + fixup launches), windows
 * labelled 1 on the per-window mma.sync kernel; results do not depend on scheduling.
 * d_hybrid_type may be NULL (all CUDA-core).  accumulate != 0 computes Y += A*X.
 * Fast paths need d_x/d_y 16-byte aligned and ldx/ldy/dim multiples of 4; anything
 * else takes a scalar kernel.                                                       */
int hcspmm_spmm(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                const int32_t *d_colidx, const int32_t *d_block_partition,
                const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                int precision, int accumulate, float *d_y, int64_t ldy, void *stream);

/* ---- A8: fused Aggregation + Update ----------------------------------------------
 * Replaces spmm_forward_plus_fixed32_fused / _final_fused / _GIN_final_fused (and
 * the _64 variants), hybrid_all_kernel.cu:596-863, 1639-2770 (pybind
 * hybrid_all.cpp:310-498) for ANY dim / hidden:  Z = A*X  (written to d_z) and
 * out = Z*W with W row-major [dim, hidden] (leading dim ldw), both operands of the
 * second product rounded to TF32, FP32 accumulate (:1809-1837).                     */
int hcspmm_spmm_gemm(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                     const int32_t *d_colidx, const int32_t *d_block_partition,
                     const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                     const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                     int precision, const float *d_w, int64_t ldw, int32_t hidden,
                     float *d_out, int64_t ldo, float *d_z, int64_t ldz, void *stream);

/* out[m, n] = a[m, k] * b[k, n], row-major FP32 in/out, TF32 tensor-core product
 * (the Update GEMM on its own).                                                      */
int hcspmm_gemm_tf32(const float *d_a, int64_t lda, const float *d_b, int64_t ldb,
                     int32_t m, int32_t k, int32_t n, float *d_out, int64_t ldo,
                     void *stream);

/* ---- dense super-window plan (tcgen05 / TMEM path) ----------------------------------------
 * Groups of eight 16-row windows (128 rows) whose windows are all labelled tensor-core are
 * multiplied as one dense contraction per group with tcgen05.mma.  The plan is
 * derived from the CSR and the window labels (it does not change any reference array):
 *   1. hcspmm_dense_plan_count   ranks columns per 128-row group, selects the dense groups
 *                                (mean column reuse >= min_reuse_x2 / 2), synchronises, and
 *                                returns h_counts = {n_dense, total condensed columns} on the host
 *   2. hcspmm_dense_plan_fill    writes the plan (hcspmm_dense_plan_words() int32 words)
 *   3. hcspmm_spmm_plan          hcspmm_spmm + plan: dense groups on the tcgen05 kernel, every other
 *                                window on the hybrid kernel.  Falls back to hcspmm_spmm when the
 *                                plan is empty, dim is not a multiple of 16 in [16, 256], precision
 *                                is not TF32, or the "umma" knob is 0.                            */
size_t hcspmm_dense_plan_workspace_bytes(int32_t n_rows, int64_t nnz);
int hcspmm_dense_plan_count(const int32_t *d_colidx, const int32_t *d_rowptr, const int32_t *d_hybrid_type,
                            int32_t n_rows, int64_t nnz, int min_reuse_x2, void *d_workspace,
                            size_t workspace_bytes, int32_t *h_counts, void *stream);
size_t hcspmm_dense_plan_words(int32_t n_rows, int32_t n_dense, int64_t total_cols);
int hcspmm_dense_plan_fill(const int32_t *d_colidx, const int32_t *d_edge_to_row, int32_t n_rows, int64_t nnz,
                           void *d_workspace, int32_t n_dense, int64_t total_cols, int32_t *d_plan,
                           size_t plan_words, void *stream);
int hcspmm_spmm_plan(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                     const int32_t *d_colidx, const int32_t *d_block_partition,
                     const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                     const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                     int precision, int accumulate, float *d_y, int64_t ldy, const int32_t *d_plan,
                     int32_t n_dense, int64_t total_cols, void *stream);

/* ---- per-graph products of a static graph (optional) ------------------------------------------------------
 * A GNN aggregates over the SAME graph at every layer and step, so what depends on the graph only is computed
 * once -- by the caller's preprocess step -- and handed to every aggregation through hcspmm_aux_t:
 *   d_splits / splits_chunk / n_splits   merge-path split points of the work-balanced kernel at diagonals
 *                                        i * splits_chunk, i = 0 .. n_splits (n_splits = ceil((n_rows + nnz) /
 *                                        splits_chunk); HCSPMM_SPLITS_CHUNK serves both default item sizes),
 *                                        from hcspmm_merge_path_splits; NULL = recomputed per call
 *   n_tc_windows                         number of windows labelled 1 (0 skips the mma.sync launch; -1 unknown)
 *   d_plan / n_dense / total_cols        dense super-window plan (hcspmm_dense_plan_fill) or NULL; plan_full = 1 when
 *                                        the plan covers every row that has stored entries (then the fused entry
 *                                        point below runs Aggregation + Update as ONE kernel)
 *   d_colidx_tagged                      the CSR's column ids with the hotness class of their column in bits 29..31
 *                                        (hcspmm_tag_columns): when X does not fit the L2 budget (knob "l2_hot_mb"),
 *                                        the balanced kernel loads the hottest rows evict_last and the others
 *                                        evict_first instead of leaving residency to chance; NULL = every row evict_last
 *   d_colidx_segments / segment_x        SEGMENT MODE: X lives in up to 8 buffers of equal row pitch; the CSR's
 *                                        column ids carry, in bits 29..31, the segment their row is read from
 *                                        (0 = d_x, s = segment_x[s]) and in bits 0..28 the row inside that segment.
 *                                        The multi-GPU layer uses it to read rows of X that a shard references only
 *                                        once or twice IN PLACE from the owners' peer-mapped exchange operands
 *                                        (NVLink loads issued by the gather itself, overlapping the sums) instead of
 *                                        copying them into the local operand first.  x_rows
 *                                        bounds the row index of every segment.  CUDA-core balanced kernel only:
 *                                        needs n_tc_windows = 0, no dense plan, 32-byte aligned rows, dim <= 256
 *                                        (wider operands: column blocks); d_colidx itself is ignored.  NULL = off
 *   d_sorted_rowptr / _colidx / _row_id  the row-sorted copy of the CSR (hcspmm_row_sort, below): the balanced kernel
 *                                        walks it instead of (d_rowptr, d_colidx) and writes sorted row i to
 *                                        Y[d_sorted_row_id[i]]; d_splits must be the split points of d_sorted_rowptr.
 *                                        Ignored in segment mode (split points are then recomputed).  NULL = off
 *   d_workspace / workspace_bytes        >= hcspmm_spmm_workspace_bytes() bytes, 16-byte aligned, for the row
 *                                        pieces of the balanced kernel; NULL = the library's private pool
 * hcspmm_spmm_aux(..., NULL, ...) is hcspmm_spmm.                                                              */
#define HCSPMM_SPLITS_CHUNK 4096
typedef struct {
  const int32_t *d_splits;
  int32_t splits_chunk, n_splits;
  int32_t n_tc_windows;
  const int32_t *d_plan;
  int32_t n_dense;
  int32_t plan_full;     /* 1: every row with stored entries lies in a dense super-window of the plan */
  int64_t total_cols;
  void *d_workspace;
  size_t workspace_bytes;
  const int32_t *d_colidx_tagged;   /* hcspmm_tag_columns output, or NULL */
  const int32_t *d_colidx_segments; /* segment mode (multi-GPU), or NULL: see below */
  const void *segment_x[8];         /* segment s = 1..7: base of that X buffer (same row pitch as d_x); [0] unused */
  const int32_t *d_sorted_rowptr;   /* hcspmm_row_sort outputs, or NULL: the balanced kernel walks the row-sorted copy */
  const int32_t *d_sorted_colidx;   /*   of the CSR and writes sorted row i to Y[d_sorted_row_id[i]]; d_splits must then */
  const int32_t *d_sorted_row_id;   /*   be the split points of d_sorted_rowptr                                          */
} hcspmm_aux_t;
/* Row-sorted copy of a CSR for the balanced kernel (csrc/rowsort.cu): rows grouped by the power of two of their length,
 * longest first, original order inside a class; every row keeps its entries in their original order, so each row sum
 * is unchanged.  Items of the balanced kernel then hold rows of similar length (products shape: -15 %).  Invisible to
 * the caller: no vertex is relabelled.  d_row_id[n_rows], d_sorted_rowptr[n_rows + 1], d_sorted_colidx[nnz].          */
size_t hcspmm_row_sort_workspace_bytes(int32_t n_rows);
int hcspmm_row_sort(const int32_t *d_rowptr, const int32_t *d_colidx, int32_t n_rows, int64_t nnz, int32_t *d_row_id,
                    int32_t *d_sorted_rowptr, int32_t *d_sorted_colidx, void *d_workspace, size_t workspace_bytes,
                    void *stream);
size_t hcspmm_tag_columns_workspace_bytes(int32_t n_cols, int64_t nnz);
int hcspmm_tag_columns(const int32_t *d_colidx, int64_t nnz, int32_t n_cols, int32_t *d_tagged, void *d_workspace,
                       size_t workspace_bytes, void *stream);
size_t hcspmm_merge_path_count(int32_t n_rows, int64_t nnz, int32_t chunk);      /* entries of d_splits: n_splits + 1 */
int hcspmm_merge_path_splits(const int32_t *d_rowptr, int32_t n_rows, int64_t nnz, int32_t chunk, int32_t *d_splits,
                             void *stream);
size_t hcspmm_spmm_workspace_bytes(int32_t n_rows, int64_t nnz, int32_t dim);
int hcspmm_spmm_aux(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                    const int32_t *d_colidx, const int32_t *d_block_partition,
                    const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                    const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                    int precision, int accumulate, float *d_y, int64_t ldy, const hcspmm_aux_t *aux, void *stream);

/* hcspmm_spmm_gemm with the graph's products.  When the dense plan covers the whole graph (plan_full), dim is a
 * multiple of 16 <= 256, hidden <= 256 and precision is TF32, Z = A X and out = Z W are ONE kernel
 * (csrc/dense_tma.cu): the aggregate of a 128-row super-window stays in tensor memory, is rounded there and is the
 * A operand of the Update tcgen05.mma -- the reference's fusion (hybrid_all_kernel.cu:1809-1837) at tcgen05 tile
 * size.  Otherwise: aggregation (plan-aware), then the TMA Update GEMM.  Knobs: "dense_tma", "fuse_update".     */
int hcspmm_spmm_gemm_aux(const float *d_x, int64_t ldx, int32_t x_rows, const int32_t *d_rowptr,
                         const int32_t *d_colidx, const int32_t *d_block_partition,
                         const int32_t *d_edge_to_column, const int32_t *d_edge_to_row,
                         const int32_t *d_hybrid_type, int32_t n_rows, int64_t nnz, int32_t dim,
                         int precision, const float *d_w, int64_t ldw, int32_t hidden,
                         float *d_out, int64_t ldo, float *d_z, int64_t ldz, const hcspmm_aux_t *aux, void *stream);

/* 1 if a tcgen05 kernel reported a barrier timeout since the last call (synchronises the device),
 * 0 if not, -1 on error.  Debug / test aid.                                              */
int hcspmm_debug_umma_error(void);

/* Measurement aid: the L2 -> SM gather roof.  Every warp of `ctas` CTAs (256 threads) sums `iters` pseudo-random
 * rows of the [rows, row_floats] FP32 buffer d_buf (row_floats 256 or 512; choose rows * row_floats * 4 well below
 * the 126 MB L2) with the SpMM's own 256-bit evict_last loads.  Bytes gathered = ctas * 8 * iters(rounded up to 8)
 * * row_floats * 4; the caller times the launch.  bench.py reports the SpMM's `roofline` against this number.  */
int hcspmm_debug_l2_gather(const float *d_buf, int32_t rows, int32_t row_floats, int32_t iters, int32_t ctas,
                           float *d_sink, void *stream);

/* ---- A9: LOA vertex reordering ------------------------------------------------------
 * Replaces reorder_plus_new_direct + the output loop of main, /root/reference/LOI.cpp:660-805,
 * 873-891 (the reference runs it offline on the CPU).  Inputs: the CSR and the CSC of the same
 * graph (the reference function takes both: row_id/col_id and row_id_in/col_id_in, :660); the
 * CSC lists must be ascending, as main builds them (:826-841); CSR rows must be sorted.
 * max_degree = largest CSR row length (sizes the scratch).  Outputs, bit-exact with the reference:
 *   d_perm[n]         vertex ids in the order main writes reorder_direct.txt: full 16-blocks in
 *                     creation order, then partial blocks, then never-visited vertices ascending
 *   d_block_start[n+1] (optional) prefix offsets of the blocks in creation order
 *   d_counts[2]       {number of blocks, number of full blocks (what main prints)}
 * The greedy is sequential across blocks by definition; the kernel is one persistent CTA that
 * parallelises the scan / arg-max / set-merge inside each of the 15 steps of a block.          */
size_t hcspmm_loa_workspace_bytes(int32_t n, int64_t nnz, int32_t max_degree);
int hcspmm_loa_reorder(const int32_t *d_rowptr, const int32_t *d_colidx, const int32_t *d_rowptr_in,
                       const int32_t *d_colidx_in, int32_t n, int64_t nnz, int32_t max_degree,
                       int32_t *d_perm, int32_t *d_block_start, int32_t *d_counts, void *d_workspace,
                       size_t workspace_bytes, void *stream);

/* ---- multi-GPU: the per-layer exchange of X over NVLink peer memory ------------------------------
 * The reference is single-GPU.  Here A is partitioned by row windows, one process per GPU, and a rank
 * needs the rows of X its shard references.  Instead of a collective every rank pulls exactly those rows
 * out of the owners' memory with NVLink loads:
 *   hcspmm_peer_alloc    cudaMalloc (zero-filled) + CUDA IPC handle (64 bytes) for a buffer peers will read
 *   hcspmm_peer_open     map a peer's buffer from its handle (another process on the same node)
 *   hcspmm_peer_close / hcspmm_peer_free
 *   hcspmm_peer_barrier  stream-ordered barrier across the ranks: d_flag_ptrs[s] is rank s's int32[world]
 *                        flag array (peer-mapped); epoch must increase by one per call on every rank.
 *                        A peer that does not arrive within "barrier_timeout_ms" sets *d_err = 1 + that peer's
 *                        rank instead of hanging; d_err may be pinned host memory (the host then sees it
 *                        without a synchronisation).  Everything computed after a timed-out barrier is
 *                        UNDEFINED: callers must check *d_err and stop.
 *   hcspmm_halo_pull     d_dst[i, col0 .. col0+width) = d_peer_x[s][d_src_row[i], col0 .. col0+width) for the
 *                        operand rows i in [d_seg[s], d_seg[s+1]) of every owner s whose bit is set in
 *                        owner_mask (the caller leaves its own bit clear: its rows are written in place).
 *                        rows = d_seg[world].  Row batches are dealt round-robin over those owners,
 *                        starting at first_owner (use own rank + 1), so all NVLink peers are read at once
 *                        and the ranks do not gang up on one owner.  width, col0, lds, ldd multiples of 4
 *                        floats.  Knob "pull_ctas" caps the grid (a pull that overlaps an SpMM should
 *                        leave it the SMs: the copy is NVLink-bound).                                    */
int hcspmm_peer_alloc(size_t bytes, void **d_ptr, void *handle64);
int hcspmm_peer_open(const void *handle64, void **d_ptr);
int hcspmm_peer_close(void *d_ptr);
int hcspmm_peer_free(void *d_ptr);
int hcspmm_peer_barrier(int32_t *const *d_flag_ptrs, int32_t rank, int32_t world, int32_t epoch, int32_t *d_err,
                        void *stream);
int hcspmm_halo_pull(const float *const *d_peer_x, int64_t lds, const int32_t *d_src_row, const int32_t *d_seg,
                     int32_t world, uint64_t owner_mask, int32_t first_owner, int32_t rows, int32_t col0, int32_t width,
                     float *d_dst, int64_t ldd, void *stream);
/* hcspmm_halo_pull for a SUBSET of the halo: list entry i (owner o: i in [d_seg[o], d_seg[o+1])) is the owner's row
 * d_src_row[i] and lands in operand row d_dst_row[i] (NULL: row i, i.e. hcspmm_halo_pull).  The row-block pipeline
 * of the multi-GPU layer pulls the halo in the order the shard's row blocks need it, block b's SpMM starting as soon
 * as its part has landed while the later parts still travel.                                  */
int hcspmm_halo_pull_rows(const float *const *d_peer_x, int64_t lds, const int32_t *d_src_row, const int32_t *d_dst_row,
                          const int32_t *d_seg, int32_t world, uint64_t owner_mask, int32_t first_owner, int32_t rows,
                          int32_t col0, int32_t width, float *d_dst, int64_t ldd, void *stream);

/* The same exchange as a PUSH by the owner: rows d_send_row[j], j in [d_send_seg[s], d_send_seg[s+1]), of d_src are
 * written to peer s's operand at d_dst_base[s] + (j - d_send_seg[s]) * ldd (d_dst_base[s] = peer s's mapped operand
 * buffer advanced to this rank's segment), for every peer whose bit is set in peer_mask.  rows = d_send_seg[world].
 * NVLink stores are posted, loads wait for a response: measured 2 GPUs, 512-byte rows, pull 470 GB/s per rank.
 * The caller's following hcspmm_peer_barrier makes the rows visible to the peers' SpMM.                          */
int hcspmm_halo_push(const float *d_src, int64_t lds, const int32_t *d_send_row, const int32_t *d_send_seg,
                     float *const *d_dst_base, int64_t ldd, int32_t world, uint64_t peer_mask, int32_t first_peer,
                     int32_t rows, int32_t col0, int32_t width, void *stream);

/* d_out[r, 0..dim) (bfloat16, row pitch ld_out elements) = round-to-nearest-even of d_x[r, 0..dim). */
int hcspmm_f32_to_bf16(const float *d_x, int64_t ldx, int32_t rows, int32_t dim, void *d_out, int64_t ld_out,
                       void *stream);

/* ---- host-buffer convenience (what a non-torch caller binds) ----------------------
 * A graph handle owns device copies of the CSR and of the preprocessing products.
 * hcspmm_graph_spmm_host copies X from (ideally pinned) host memory, runs the
 * kernel, and copies Y back: host -> device -> host inside the call.                 */
typedef struct hcspmm_graph hcspmm_graph_t;
int hcspmm_graph_create(const int32_t *h_rowptr, const int32_t *h_colidx, int32_t n_rows,
                        int64_t nnz, int32_t x_rows, int classifier, hcspmm_graph_t **out);
int hcspmm_graph_spmm_host(hcspmm_graph_t *g, const float *h_x, int32_t dim, int precision,
                           float *h_y);
/* copies the four preprocessing arrays to host buffers (any of them may be NULL)      */
int hcspmm_graph_get_preprocess(hcspmm_graph_t *g, int32_t *h_block_partition,
                                int32_t *h_edge_to_column, int32_t *h_edge_to_row,
                                int32_t *h_hybrid_type);
void hcspmm_graph_destroy(hcspmm_graph_t *g);

#ifdef __cplusplus
}
#endif
#endif /* HCSPMM_H_ */
