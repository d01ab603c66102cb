// torch_shim.cpp -- the `HCSPMM` Python extension module: a thin torch layer over the C ABI
// of libhcspmm.so (include/hcspmm.h).
//
// Mirrors the reference's pybind module, /root/reference/hybrid_kernel/hybrid_all.cpp:500-525:
// the same 18 names with the same positional signatures and list-of-tensor returns, so
// /root/reference/GNN_model.py and HC-SpMM_main.py import and call it unchanged.  Differences,
// all in the direction of "works where the reference does not":
//   * every entry point accepts any dim / hidden (the reference needs 32 / 64 / <= 48);
//   * inputs are checked for device, dtype, shape and contiguity (the reference checks
//     "is CUDA" and "is contiguous" on the first six tensors only, hybrid_all.cpp:185-210);
//   * `weights` may be a strided view -- it is made contiguous (the reference reads its raw
//     memory, hybrid_all_kernel.cu:645,753; set_bug_compat(True) restores that reading);
//   * launches go to the CURRENT stream of the input's device, outputs come from the caching
//     allocator (the reference uses the legacy stream and leaks cudaMalloc blobs, :356-372);
//   * the operator may be rectangular: output rows = nodePointer.size(0) - 1, X rows =
//     input.size(0) (row-partitioned shards keep global column ids).
// There is no CPU path: a non-CUDA tensor is an error.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <string.h>

#include <algorithm>

#include <vector>

#include "../../include/hcspmm.h"

namespace {

// Per-graph products travel in the two tensors the reference returns as one-element placeholders
// (hybrid_all_kernel.cu:405) and that its callers only store and pass back:
//   row_nzr  HOST int32[kHdrWords]  header (so that forward() needs no device synchronisation to read it)
//   col_nzr  DEVICE int32           [dense plan words | merge-path split points]
// The selector / precision / dense settings in force when preprocess() ran are recorded in the header and are
// what forward*() uses for THAT graph: set_classifier / set_precision / set_dense only set the defaults for
// graphs preprocessed afterwards, so two graphs (or two threads) never see each other's settings.
constexpr int32_t kPlanMagic = 0x48435044;
enum { H_MAGIC = 0, H_NDENSE, H_TOTALCOLS, H_NROWS, H_VERSION, H_PRECISION, H_DENSE, H_NTC, H_SPLITS_CHUNK, H_NSPLITS,
       H_SPLITS_OFF, H_CLASSIFIER, H_PLAN_FULL, H_TAG_OFF, H_NCOLS, H_SORT_OFF, kHdrWords = 16 };
bool g_dense = false;                        // defaults for the next preprocess()
int g_classifier = HCSPMM_CLASSIFIER_SHIPPED;
int g_precision = HCSPMM_PRECISION_TF32;
bool g_bug_compat = false;
bool g_row_sort = true;                      // preprocess() keeps a row-sorted copy of low-degree CSRs for the balanced
                                             // kernel (hcspmm_row_sort; invisible to the caller)
bool g_tag_columns = false;                  // preprocess() also emits hotness-tagged column ids (L2 residency hints:
                                             // they pay from 2 KB rows = dim 512 upwards, so they are opt-in)

void check_rc(int rc, const char *what) {
  TORCH_CHECK(rc == 0, "HCSPMM.", what, " failed (code ", rc, "): ", hcspmm_last_error());
}

void check_i32(const torch::Tensor &t, const char *name, const torch::Tensor &like) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(t.scalar_type() == torch::kInt32, name, " must be int32");
  TORCH_CHECK(t.device() == like.device(), name, " must be on the same device as the input");
}

void check_f32_2d(const torch::Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
  TORCH_CHECK(t.dim() == 2, name, " must be 2-D");
}

struct Graph {
  const int32_t *rowptr, *colidx, *bp, *etc, *etr, *ht;
  int32_t n_rows;
  int64_t nnz;
};

Graph check_graph(const torch::Tensor &input, const torch::Tensor &nodePointer,
                  const torch::Tensor &edgeList, const torch::Tensor &blockPartition,
                  const torch::Tensor &edgeToColumn, const torch::Tensor &edgeToRow,
                  const torch::Tensor &hybrid_type, bool strided_input = false) {
  if (!strided_input) check_f32_2d(input, "input");
  check_i32(nodePointer, "nodePointer", input);
  check_i32(edgeList, "edgeList", input);
  check_i32(blockPartition, "blockPartition", input);
  check_i32(edgeToColumn, "edgeToColumn", input);
  check_i32(edgeToRow, "edgeToRow", input);
  check_i32(hybrid_type, "hybrid_type", input);
  TORCH_CHECK(nodePointer.numel() >= 1, "nodePointer must have num_nodes + 1 entries");
  Graph g;
  g.n_rows = (int32_t)(nodePointer.size(0) - 1);  // hybrid_all.cpp:212
  g.nnz = edgeList.size(0);                       // hybrid_all.cpp:213
  const int64_t w = (g.n_rows + HCSPMM_BLK_H - 1) / HCSPMM_BLK_H;
  TORCH_CHECK(blockPartition.numel() >= w && hybrid_type.numel() >= w,
              "blockPartition / hybrid_type must have one entry per 16-row window");
  TORCH_CHECK(edgeToColumn.numel() >= g.nnz && edgeToRow.numel() >= g.nnz,
              "edgeToColumn / edgeToRow must have one entry per edge");
  g.rowptr = nodePointer.data_ptr<int32_t>();
  g.colidx = edgeList.data_ptr<int32_t>();
  g.bp = blockPartition.data_ptr<int32_t>();
  g.etc = edgeToColumn.data_ptr<int32_t>();
  g.etr = edgeToRow.data_ptr<int32_t>();
  g.ht = hybrid_type.data_ptr<int32_t>();
  return g;
}

// preprocess, hybrid_all.cpp:13-17 / :501 -> hybrid_all_kernel.cu:339-408
std::vector<torch::Tensor> preprocess(torch::Tensor edgeList, torch::Tensor nodePointer,
                                      int64_t num_nodes, int64_t edge_num, int64_t num_row_windows) {
  TORCH_CHECK(edgeList.is_cuda() && nodePointer.is_cuda(), "preprocess: tensors must be CUDA tensors");
  TORCH_CHECK(edgeList.scalar_type() == torch::kInt32 && nodePointer.scalar_type() == torch::kInt32,
              "preprocess: edgeList / nodePointer must be int32");
  TORCH_CHECK(edgeList.is_contiguous() && nodePointer.is_contiguous(), "preprocess: tensors must be contiguous");
  TORCH_CHECK(nodePointer.numel() == num_nodes + 1, "preprocess: nodePointer must have num_nodes + 1 entries");
  TORCH_CHECK(edgeList.numel() == edge_num, "preprocess: edgeList must have edge_num entries");
  TORCH_CHECK(num_row_windows == (num_nodes + HCSPMM_BLK_H - 1) / HCSPMM_BLK_H,
              "preprocess: num_row_windows must be ceil(num_nodes / 16)");
  c10::cuda::CUDAGuard guard(edgeList.device());
  auto opts = torch::TensorOptions().dtype(torch::kInt32).device(edgeList.device());
  auto bp = torch::zeros({num_row_windows}, opts);
  auto ht = torch::zeros({num_row_windows}, opts);
  auto etc = torch::zeros({edge_num}, opts);
  auto etr = torch::zeros({edge_num}, opts);
  const size_t ws_bytes = hcspmm_preprocess_workspace_bytes((int32_t)num_nodes, edge_num);
  auto ws = torch::empty({(int64_t)ws_bytes}, opts.dtype(torch::kUInt8));
  check_rc(hcspmm_preprocess(edgeList.data_ptr<int32_t>(), nodePointer.data_ptr<int32_t>(),
                             (int32_t)num_nodes, edge_num, (int32_t)num_row_windows, g_classifier,
                             bp.data_ptr<int32_t>(), etc.data_ptr<int32_t>(), etr.data_ptr<int32_t>(),
                             ht.data_ptr<int32_t>(), ws.data_ptr(), ws_bytes,
                             at::cuda::getCurrentCUDAStream().stream()),
           "preprocess");
  // header + per-graph products (see the top of this file)
  auto stream = at::cuda::getCurrentCUDAStream().stream();
  auto row_nzr = torch::zeros({kHdrWords}, torch::TensorOptions().dtype(torch::kInt32));
  int32_t *h = row_nzr.data_ptr<int32_t>();
  h[H_MAGIC] = kPlanMagic; h[H_NROWS] = (int32_t)num_nodes; h[H_VERSION] = 2; h[H_PRECISION] = g_precision;
  h[H_DENSE] = g_dense ? 1 : 0; h[H_CLASSIFIER] = g_classifier;
  h[H_NTC] = edge_num > 0 ? (int32_t)ht.eq(1).sum().item<int64_t>() : 0;
  torch::Tensor plan;
  if (g_dense && edge_num > 0 && g_classifier != HCSPMM_CLASSIFIER_SHIPPED && g_classifier != HCSPMM_CLASSIFIER_ALL_CUDA) {
    const size_t pws = hcspmm_dense_plan_workspace_bytes((int32_t)num_nodes, edge_num);
    auto pw = torch::empty({(int64_t)pws}, opts.dtype(torch::kUInt8));
    int32_t counts[2] = {0, 0};
    const int min_reuse_x2 = g_classifier == HCSPMM_CLASSIFIER_ALL_TC ? 0 : 4;
    check_rc(hcspmm_dense_plan_count(edgeList.data_ptr<int32_t>(), nodePointer.data_ptr<int32_t>(), ht.data_ptr<int32_t>(),
                                     (int32_t)num_nodes, edge_num, min_reuse_x2, pw.data_ptr(), pws, counts, stream),
             "dense_plan_count");
    if (counts[0] > 0) {
      const size_t words = hcspmm_dense_plan_words((int32_t)num_nodes, counts[0], counts[1]);
      plan = torch::zeros({(int64_t)words}, opts);
      check_rc(hcspmm_dense_plan_fill(edgeList.data_ptr<int32_t>(), etr.data_ptr<int32_t>(), (int32_t)num_nodes, edge_num,
                                      pw.data_ptr(), counts[0], counts[1], plan.data_ptr<int32_t>(), words, stream),
               "dense_plan_fill");
      h[H_NDENSE] = counts[0]; h[H_TOTALCOLS] = counts[1];
      // does the plan cover every 128-row super-window that has entries?  (then *_fused is one kernel)
      auto edges = torch::arange(0, num_nodes + 128, 128, opts.dtype(torch::kInt64)).clamp_max(num_nodes);
      auto rp64 = nodePointer.index_select(0, edges);
      const int64_t busy = rp64.slice(0, 1).gt(rp64.slice(0, 0, -1)).sum().item<int64_t>();
      h[H_PLAN_FULL] = busy == counts[0] ? 1 : 0;
    }
  }
  // merge-path split points of the work-balanced kernel (static per graph)
  const int64_t plan_words = plan.defined() ? plan.numel() : 0;
  const int64_t n_split_words = ((int64_t)hcspmm_merge_path_count((int32_t)num_nodes, edge_num, HCSPMM_SPLITS_CHUNK) + 3) / 4 * 4;
  // hotness-tagged column ids for the balanced kernel's L2 residency hints: worth their nnz words only when a
  // gathered matrix of this many rows can exceed L2 at all (>= 128-byte rows against a 64 MB budget)
  int64_t n_cols = 0, tag_words = 0;
  if (g_tag_columns && edge_num >= 8 * num_nodes && edge_num > 0) {
    n_cols = std::max<int64_t>(num_nodes, edgeList.max().item<int64_t>() + 1);
    if (n_cols * 128 > (64LL << 20) && n_cols < (1LL << 29)) tag_words = edge_num;
  }
  // row-sorted copy of the CSR (rows grouped by the power of two of their length): low-degree graphs on the balanced
  // kernel -- an item then holds rows of similar length (products shape: -15 %).  [row_id | rowptr_s | colidx_s]
  const bool row_sort = g_row_sort && tag_words == 0 && edge_num >= 8 * num_nodes && edge_num < 64 * num_nodes && num_nodes > 0;
  const int64_t n4 = (num_nodes + 3) / 4 * 4, n14 = (num_nodes + 1 + 3) / 4 * 4;
  const int64_t sort_words = row_sort ? n4 + n14 + edge_num : 0;
  TORCH_CHECK(plan_words + n_split_words + tag_words + sort_words < (1LL << 31), "preprocess: per-graph products too large");
  auto col_nzr = torch::empty({plan_words + n_split_words + tag_words + sort_words}, opts);
  const int32_t *rowptr_for_splits = nodePointer.data_ptr<int32_t>();
  if (row_sort) {
    const int64_t off = plan_words + n_split_words + tag_words;
    int32_t *base = col_nzr.data_ptr<int32_t>() + off;
    const size_t sws = hcspmm_row_sort_workspace_bytes((int32_t)num_nodes);
    auto sw = torch::empty({(int64_t)sws}, opts.dtype(torch::kUInt8));
    check_rc(hcspmm_row_sort(nodePointer.data_ptr<int32_t>(), edgeList.data_ptr<int32_t>(), (int32_t)num_nodes, edge_num,
                             base, base + n4, base + n4 + n14, sw.data_ptr(), sws, stream),
             "row_sort");
    h[H_SORT_OFF] = (int32_t)off;
    rowptr_for_splits = base + n4;          // the split points below are those of the sorted copy
  }
  if (tag_words > 0) {
    const size_t tws = hcspmm_tag_columns_workspace_bytes((int32_t)n_cols, edge_num);
    auto tw = torch::empty({(int64_t)tws}, opts.dtype(torch::kUInt8));
    check_rc(hcspmm_tag_columns(edgeList.data_ptr<int32_t>(), edge_num, (int32_t)n_cols,
                                col_nzr.data_ptr<int32_t>() + plan_words + n_split_words, tw.data_ptr(), tws, stream),
             "tag_columns");
    h[H_TAG_OFF] = (int32_t)(plan_words + n_split_words); h[H_NCOLS] = (int32_t)n_cols;
  }
  if (plan_words > 0) col_nzr.narrow(0, 0, plan_words).copy_(plan);
  check_rc(hcspmm_merge_path_splits(rowptr_for_splits, (int32_t)num_nodes, edge_num, HCSPMM_SPLITS_CHUNK,
                                    col_nzr.data_ptr<int32_t>() + plan_words, stream),
           "merge_path_splits");
  h[H_SPLITS_CHUNK] = HCSPMM_SPLITS_CHUNK;
  h[H_NSPLITS] = (int32_t)hcspmm_merge_path_count((int32_t)num_nodes, edge_num, HCSPMM_SPLITS_CHUNK) - 1;
  h[H_SPLITS_OFF] = (int32_t)plan_words;
  return {bp, etc, etr, ht, row_nzr, col_nzr};
}

// the graph's own products, when row_nzr / col_nzr carry them (header written by preprocess())
struct AuxView {
  bool ok = false;
  hcspmm_aux_t aux;
  int precision = HCSPMM_PRECISION_TF32;
  torch::Tensor ws;   // keeps the workspace alive for the duration of the call
};

AuxView make_aux(const torch::Tensor &input, const Graph &g, const torch::Tensor &row_nzr, const torch::Tensor &col_nzr,
                 int64_t dim) {
  AuxView v;
  memset(&v.aux, 0, sizeof(v.aux));
  v.precision = g_precision;
  if (!(row_nzr.defined() && row_nzr.device().is_cpu() && row_nzr.scalar_type() == torch::kInt32 &&
        row_nzr.numel() >= kHdrWords && row_nzr.data_ptr<int32_t>()[H_MAGIC] == kPlanMagic &&
        row_nzr.data_ptr<int32_t>()[H_VERSION] == 2 && row_nzr.data_ptr<int32_t>()[H_NROWS] == g.n_rows &&
        col_nzr.defined() && col_nzr.is_cuda() && col_nzr.scalar_type() == torch::kInt32 && col_nzr.is_contiguous() &&
        col_nzr.device() == input.device()))
    return v;
  const int32_t *h = row_nzr.data_ptr<int32_t>();
  const int32_t *blob = col_nzr.data_ptr<int32_t>();
  v.ok = true;
  v.precision = h[H_PRECISION];
  v.aux.n_tc_windows = h[H_NTC];
  if (h[H_NSPLITS] > 0 && col_nzr.numel() >= (int64_t)h[H_SPLITS_OFF] + h[H_NSPLITS] + 1) {
    v.aux.d_splits = blob + h[H_SPLITS_OFF]; v.aux.splits_chunk = h[H_SPLITS_CHUNK]; v.aux.n_splits = h[H_NSPLITS];
  }
  if (h[H_TAG_OFF] > 0 && col_nzr.numel() >= (int64_t)h[H_TAG_OFF] + g.nnz) v.aux.d_colidx_tagged = blob + h[H_TAG_OFF];
  if (h[H_SORT_OFF] > 0) {
    const int64_t n4 = ((int64_t)g.n_rows + 3) / 4 * 4, n14 = ((int64_t)g.n_rows + 1 + 3) / 4 * 4;
    if (col_nzr.numel() >= (int64_t)h[H_SORT_OFF] + n4 + n14 + g.nnz) {
      v.aux.d_sorted_row_id = blob + h[H_SORT_OFF];
      v.aux.d_sorted_rowptr = blob + h[H_SORT_OFF] + n4;
      v.aux.d_sorted_colidx = blob + h[H_SORT_OFF] + n4 + n14;
    } else {
      v.aux.d_splits = nullptr;           // cannot happen with tensors from preprocess(); never pair foreign splits
    }
  }
  if (h[H_DENSE] && h[H_NDENSE] > 0) {
    v.aux.d_plan = blob; v.aux.n_dense = h[H_NDENSE]; v.aux.total_cols = h[H_TOTALCOLS]; v.aux.plan_full = h[H_PLAN_FULL];
  }
  if (g.nnz >= 8LL * g.n_rows) {   // the balanced kernel's row pieces, from the caching allocator
    const size_t ws_bytes = hcspmm_spmm_workspace_bytes(g.n_rows, g.nnz, (int32_t)dim);
    v.ws = torch::empty({(int64_t)ws_bytes}, input.options().dtype(torch::kUInt8));
    v.aux.d_workspace = v.ws.data_ptr(); v.aux.workspace_bytes = ws_bytes;
  }
  return v;
}

// one aggregation through the C ABI
void run_spmm(const torch::Tensor &input, const Graph &g, const torch::Tensor &row_nzr, const torch::Tensor &col_nzr,
              float *out, int64_t ldy, int accumulate, const char *what, bool bf16_stored = false) {
  const int64_t dim = input.size(1);
  auto stream = at::cuda::getCurrentCUDAStream().stream();
  // a BF16-stored operand (multi-GPU exchange buffer) is passed as its raw rows; ldx counts bfloat16 elements
  const float *xptr = reinterpret_cast<const float *>(input.data_ptr());
  AuxView v = make_aux(input, g, row_nzr, col_nzr, dim);
  check_rc(hcspmm_spmm_aux(xptr, input.stride(0), (int32_t)input.size(0), g.rowptr, g.colidx, g.bp, g.etc, g.etr, g.ht,
                           g.n_rows, g.nnz, (int32_t)dim, bf16_stored ? HCSPMM_PRECISION_BF16_STORED : v.precision,
                           accumulate, out, ldy, v.ok ? &v.aux : nullptr, stream),
           what);
}

// forward / forward_more / forward_fixed32 / forward_fixed64 and their backward_* aliases,
// hybrid_all.cpp:194-308 -> hybrid_all_kernel.cu:410-595
std::vector<torch::Tensor> spmm_forward(torch::Tensor input, torch::Tensor nodePointer,
                                        torch::Tensor edgeList, torch::Tensor blockPartition,
                                        torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                                        torch::Tensor hybrid_type, torch::Tensor row_nzr,
                                        torch::Tensor col_nzr) {
  Graph g = check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, hybrid_type);
  c10::cuda::CUDAGuard guard(input.device());
  const int64_t dim = input.size(1);
  auto out = torch::empty({g.n_rows, dim}, input.options());
  run_spmm(input, g, row_nzr, col_nzr, out.data_ptr<float>(), dim, 0, "forward");
  return {out};
}

// out (+)= A * X with row-strided X / out views (stride(1) == 1): the kernel takes ldx / ldy.  Used by
// the multi-GPU layer (feature-slab pipelining, shard accumulation).
torch::Tensor spmm_strided(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                           torch::Tensor blockPartition, torch::Tensor edgeToColumn,
                           torch::Tensor edgeToRow, torch::Tensor hybrid_type, torch::Tensor out,
                           bool accumulate, c10::optional<torch::Tensor> row_nzr, c10::optional<torch::Tensor> col_nzr) {
  TORCH_CHECK(input.is_cuda() && input.scalar_type() == torch::kFloat32 && input.dim() == 2 &&
              (input.stride(1) == 1 || input.size(1) == 1), "input must be a 2-D float32 CUDA tensor with unit column stride");
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == torch::kFloat32 && out.dim() == 2 &&
              (out.stride(1) == 1 || out.size(1) == 1), "out must be a 2-D float32 CUDA tensor with unit column stride");
  TORCH_CHECK(out.device() == input.device(), "out must be on the input's device");
  Graph g = check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, hybrid_type, true);
  TORCH_CHECK(out.size(0) == g.n_rows && out.size(1) == input.size(1), "out must be [num_nodes, dim]");
  c10::cuda::CUDAGuard guard(input.device());
  const int64_t dim = input.size(1);
  (void)dim;
  run_spmm(input, g, row_nzr.has_value() ? *row_nzr : torch::Tensor(), col_nzr.has_value() ? *col_nzr : torch::Tensor(),
           out.data_ptr<float>(), out.stride(0), accumulate ? 1 : 0, "spmm_strided");
  return out;
}

// out (+)= A * X for an X ALREADY stored as bfloat16 (the multi-GPU exchange operand whose halo rows travelled at
// half the bytes): FP32 accumulate on the CUDA-core path, no conversion pass.
torch::Tensor spmm_bf16(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor edgeList,
                        torch::Tensor blockPartition, torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                        torch::Tensor hybrid_type, torch::Tensor out, bool accumulate,
                        c10::optional<torch::Tensor> row_nzr, c10::optional<torch::Tensor> col_nzr) {
  TORCH_CHECK(input.is_cuda() && input.scalar_type() == torch::kBFloat16 && input.dim() == 2 && input.stride(1) == 1,
              "input must be a 2-D bfloat16 CUDA tensor with unit column stride");
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == torch::kFloat32 && out.dim() == 2 && out.stride(1) == 1 &&
              out.device() == input.device(), "out must be a 2-D float32 CUDA tensor on the input's device");
  Graph g = check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, hybrid_type, true);
  TORCH_CHECK(out.size(0) == g.n_rows && out.size(1) == input.size(1), "out must be [num_nodes, dim]");
  c10::cuda::CUDAGuard guard(input.device());
  run_spmm(input, g, row_nzr.has_value() ? *row_nzr : torch::Tensor(), col_nzr.has_value() ? *col_nzr : torch::Tensor(),
           out.data_ptr<float>(), out.stride(0), accumulate ? 1 : 0, "spmm_bf16", true);
  return out;
}

// out (+)= A * X with X in SEGMENTS (hcspmm_aux_t.d_colidx_segments): `input` is segment 0 (FP32 or bfloat16 rows),
// segment_ptrs[s], s = 1..7, are device addresses of further buffers with the same row pitch -- the peers' exchange
// operands, mapped over NVLink -- and colidx_segments carries the segment of every entry in bits 29..31.  x_rows
// bounds the row index inside any segment.  CUDA-core balanced kernel; row_nzr / col_nzr are the products of
// preprocess() on the shard (split points, workspace).
torch::Tensor spmm_segments(torch::Tensor input, torch::Tensor nodePointer, torch::Tensor colidx_segments,
                            std::vector<int64_t> segment_ptrs, int64_t x_rows, torch::Tensor out, bool accumulate,
                            c10::optional<torch::Tensor> row_nzr, c10::optional<torch::Tensor> col_nzr) {
  const bool b16 = input.scalar_type() == torch::kBFloat16;
  TORCH_CHECK(input.is_cuda() && (b16 || input.scalar_type() == torch::kFloat32) && input.dim() == 2 && input.stride(1) == 1,
              "input must be a 2-D float32 / bfloat16 CUDA tensor with unit column stride");
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == torch::kFloat32 && out.dim() == 2 && out.stride(1) == 1 &&
              out.device() == input.device(), "out must be a 2-D float32 CUDA tensor on the input's device");
  check_i32(nodePointer, "nodePointer", input);
  check_i32(colidx_segments, "colidx_segments", input);
  TORCH_CHECK(segment_ptrs.size() <= 8, "at most 8 segments");
  Graph g;
  g.rowptr = nodePointer.data_ptr<int32_t>();
  g.colidx = colidx_segments.data_ptr<int32_t>();
  g.bp = g.etc = g.etr = g.ht = nullptr;
  g.n_rows = (int32_t)(nodePointer.size(0) - 1);
  g.nnz = colidx_segments.size(0);
  TORCH_CHECK(out.size(0) == g.n_rows && out.size(1) == input.size(1), "out must be [num_nodes, dim]");
  TORCH_CHECK(x_rows >= input.size(0) && x_rows < (1LL << 29), "x_rows must bound every segment's rows (< 2^29)");
  c10::cuda::CUDAGuard guard(input.device());
  const int64_t dim = input.size(1);
  AuxView v = make_aux(input, g, row_nzr.has_value() ? *row_nzr : torch::Tensor(),
                       col_nzr.has_value() ? *col_nzr : torch::Tensor(), dim);
  if (!v.ok) v.aux.n_tc_windows = 0;
  v.aux.d_colidx_tagged = nullptr;
  v.aux.d_colidx_segments = g.colidx;
  for (size_t i = 0; i < segment_ptrs.size(); ++i) v.aux.segment_x[i] = reinterpret_cast<const void *>(segment_ptrs[i]);
  check_rc(hcspmm_spmm_aux(reinterpret_cast<const float *>(input.data_ptr()), input.stride(0), (int32_t)x_rows, g.rowptr,
                           g.colidx, nullptr, nullptr, nullptr, nullptr, g.n_rows, g.nnz, (int32_t)dim,
                           b16 ? HCSPMM_PRECISION_BF16_STORED : HCSPMM_PRECISION_FP32, accumulate ? 1 : 0,
                           out.data_ptr<float>(), out.stride(0), &v.aux, at::cuda::getCurrentCUDAStream().stream()),
           "spmm_segments");
  return out;
}

// out[r, :] (bfloat16 view, any row pitch) = round-to-nearest-even of input[r, :]
torch::Tensor f32_to_bf16_into(torch::Tensor input, torch::Tensor out) {
  TORCH_CHECK(input.is_cuda() && input.scalar_type() == torch::kFloat32 && input.dim() == 2 && input.stride(1) == 1,
              "input must be a 2-D float32 CUDA tensor with unit column stride");
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == torch::kBFloat16 && out.dim() == 2 && out.stride(1) == 1 &&
              out.device() == input.device() && out.size(0) == input.size(0) && out.size(1) == input.size(1),
              "out must be a bfloat16 CUDA tensor of the input's shape");
  c10::cuda::CUDAGuard guard(input.device());
  check_rc(hcspmm_f32_to_bf16(input.data_ptr<float>(), input.stride(0), (int32_t)input.size(0), (int32_t)input.size(1),
                              out.data_ptr(), out.stride(0), at::cuda::getCurrentCUDAStream().stream()),
           "f32_to_bf16");
  return out;
}

torch::Tensor dense_weights(const torch::Tensor &weights, const torch::Tensor &input) {
  TORCH_CHECK(weights.is_cuda() && weights.device() == input.device(), "weights must be on the input's device");
  TORCH_CHECK(weights.scalar_type() == torch::kFloat32 && weights.dim() == 2, "weights must be 2-D float32");
  TORCH_CHECK(weights.size(0) == input.size(1), "weights must be [dim, hidden]");
  if (g_bug_compat && !weights.is_contiguous())
    // what the reference multiplies by: the view's raw memory read as row-major [dim, hidden]
    return weights.as_strided({weights.size(0), weights.size(1)}, {weights.size(1), 1});
  return weights.contiguous();
}

std::vector<torch::Tensor> fused_impl(const torch::Tensor &input, const Graph &g, const torch::Tensor &row_nzr,
                                      const torch::Tensor &col_nzr, const torch::Tensor &weights, torch::Tensor out,
                                      const char *what) {
  c10::cuda::CUDAGuard guard(input.device());
  auto w = dense_weights(weights, input);
  const int64_t dim = input.size(1), hidden = w.size(1);
  auto z = torch::empty({g.n_rows, dim}, input.options());
  // Z = A X and out = Z W: ONE kernel when the graph's dense plan covers it (hcspmm_spmm_gemm_aux), otherwise the
  // plan-aware aggregation followed by the TMA Update GEMM
  AuxView v = make_aux(input, g, row_nzr, col_nzr, dim);
  check_rc(hcspmm_spmm_gemm_aux(input.data_ptr<float>(), input.stride(0), (int32_t)input.size(0), g.rowptr, g.colidx, g.bp,
                                g.etc, g.etr, g.ht, g.n_rows, g.nnz, (int32_t)dim, v.precision, w.data_ptr<float>(), hidden,
                                (int32_t)hidden, out.data_ptr<float>(), out.stride(0), z.data_ptr<float>(), dim,
                                v.ok ? &v.aux : nullptr, at::cuda::getCurrentCUDAStream().stream()),
           what);
  return {out, z};
}

// forward_fixed32_fused / forward_fixed64_fused / forward_GIN_final_fused (+ backward_*),
// hybrid_all.cpp:310-403, 469-498 -> hybrid_all_kernel.cu:596-700, 810-863
std::vector<torch::Tensor> spmm_forward_fused(torch::Tensor input, torch::Tensor nodePointer,
                                              torch::Tensor edgeList, torch::Tensor blockPartition,
                                              torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                                              torch::Tensor hybrid_type, torch::Tensor row_nzr,
                                              torch::Tensor col_nzr, torch::Tensor weights) {
  Graph g = check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, hybrid_type);
  TORCH_CHECK(weights.dim() == 2, "weights must be 2-D");
  auto out = torch::empty({g.n_rows, weights.size(1)}, input.options());
  return fused_impl(input, g, row_nzr, col_nzr, weights, out, "forward_fused");
}

// forward_final_fused / forward_final_fused_64 (+ backward_*): writes the caller's `output`,
// hybrid_all.cpp:405-467 -> hybrid_all_kernel.cu:701-809
std::vector<torch::Tensor> spmm_forward_final_fused(torch::Tensor input, torch::Tensor nodePointer,
                                                    torch::Tensor edgeList, torch::Tensor blockPartition,
                                                    torch::Tensor edgeToColumn, torch::Tensor edgeToRow,
                                                    torch::Tensor hybrid_type, torch::Tensor row_nzr,
                                                    torch::Tensor col_nzr, torch::Tensor weights,
                                                    torch::Tensor output) {
  Graph g = check_graph(input, nodePointer, edgeList, blockPartition, edgeToColumn, edgeToRow, hybrid_type);
  check_f32_2d(output, "output");
  TORCH_CHECK(weights.dim() == 2, "weights must be 2-D");
  TORCH_CHECK(output.device() == input.device(), "output must be on the input's device");
  TORCH_CHECK(output.size(0) == g.n_rows && output.size(1) == weights.size(1),
              "output must be [num_nodes, hidden] = [", g.n_rows, ", ", weights.size(1), "]");
  return fused_impl(input, g, row_nzr, col_nzr, weights, output, "forward_final_fused");
}

// a @ b (TF32 product, FP32 accumulate).  `out` (optional): a [m, n] float32 view with unit column stride -- e.g. the
// own-rows segment of a multi-GPU exchange operand, so the Update GEMM's result needs no staging copy.
torch::Tensor gemm_tf32(torch::Tensor a, torch::Tensor b, c10::optional<torch::Tensor> out_opt) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == torch::kFloat32 && a.dim() == 2 && (a.stride(1) == 1 || a.size(1) == 1),
              "a must be a 2-D float32 CUDA tensor with unit column stride");
  check_f32_2d(b, "b");
  TORCH_CHECK(a.device() == b.device() && a.size(1) == b.size(0), "gemm_tf32: shape/device mismatch");
  c10::cuda::CUDAGuard guard(a.device());
  torch::Tensor out;
  if (out_opt.has_value()) {
    out = *out_opt;
    TORCH_CHECK(out.is_cuda() && out.device() == a.device() && out.scalar_type() == torch::kFloat32 && out.dim() == 2 &&
                out.size(0) == a.size(0) && out.size(1) == b.size(1) && (out.stride(1) == 1 || out.size(1) == 1),
                "gemm_tf32: out must be a [m, n] float32 view with unit column stride on a's device");
  } else {
    out = torch::empty({a.size(0), b.size(1)}, a.options());
  }
  check_rc(hcspmm_gemm_tf32(a.data_ptr<float>(), a.stride(0), b.data_ptr<float>(), b.size(1), (int32_t)a.size(0),
                            (int32_t)a.size(1), (int32_t)b.size(1), out.data_ptr<float>(), out.stride(0),
                            at::cuda::getCurrentCUDAStream().stream()),
           "gemm_tf32");
  return out;
}

std::string set_classifier(const std::string &mode) {
  static const char *names[] = {"shipped", "intended", "b200", "all_cuda", "all_tc", "b200_window"};
  std::string old = names[g_classifier];
  for (int i = 0; i < 6; ++i)
    if (mode == names[i]) { g_classifier = i; return old; }
  TORCH_CHECK(false, "unknown classifier '", mode, "' (shipped|intended|b200|all_cuda|all_tc|b200_window)");
}

std::string set_precision(const std::string &mode) {
  static const char *names[] = {"tf32", "tf32x2", "fp32", "bf16"};
  std::string old = names[g_precision];
  for (int i = 0; i < 4; ++i)
    if (mode == names[i]) { g_precision = i; return old; }
  TORCH_CHECK(false, "unknown precision '", mode, "' (tf32|tf32x2|fp32|bf16)");
}

bool set_dense(bool on) {
  bool old = g_dense;
  g_dense = on;
  return old;
}

bool set_row_sort(bool on) {
  bool old = g_row_sort;
  g_row_sort = on;
  return old;
}

bool set_tag_columns(bool on) {
  bool old = g_tag_columns;
  g_tag_columns = on;
  return old;
}

bool set_bug_compat(bool on) {
  bool old = g_bug_compat;
  g_bug_compat = on;
  return old;
}

int64_t set_tuning(const std::string &key, int64_t value) {
  int old = hcspmm_set_tuning(key.c_str(), (int)value);
  TORCH_CHECK(old != -1, "unknown tuning key '", key, "'");
  return old;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "HCSPMM: B200-native hybrid SpMM (drop-in for the HC-SpMM reference extension)";
  m.def("preprocess", &preprocess, "Preprocess Step (GPU)");
  // forward computation -- names of hybrid_all.cpp:503-512
  m.def("forward", &spmm_forward, "HCSPMM SPMM forward (CUDA)");
  m.def("forward_more", &spmm_forward, "HCSPMM SPMM forward more (CUDA)");
  m.def("forward_fixed32", &spmm_forward, "HCSPMM SPMM forward fixed32 (CUDA)");
  m.def("forward_fixed32_fused", &spmm_forward_fused, "HCSPMM SPMM forward fixed32 fused (CUDA)");
  m.def("forward_final_fused", &spmm_forward_final_fused, "HCSPMM SPMM forward final fused (CUDA)");
  m.def("forward_fixed64", &spmm_forward, "HCSPMM SPMM forward fixed64 (CUDA)");
  m.def("forward_fixed64_fused", &spmm_forward_fused, "HCSPMM SPMM forward fixed64 fused (CUDA)");
  m.def("forward_final_fused_64", &spmm_forward_final_fused, "HCSPMM SPMM forward final fused 64 (CUDA)");
  m.def("forward_GIN_final_fused", &spmm_forward_fused, "HCSPMM SPMM forward for GIN final fused (CUDA)");
  // backward -- names of hybrid_all.cpp:516-523 (same functions, as in the reference)
  m.def("backward", &spmm_forward, "HCSPMM SPMM backward (CUDA)");
  m.def("backward_fixed32", &spmm_forward, "HCSPMM SPMM backward fixed32 (CUDA)");
  m.def("backward_fixed32_fused", &spmm_forward_fused, "HCSPMM SPMM backward fixed32 fused (CUDA)");
  m.def("backward_final_fused", &spmm_forward_final_fused, "HCSPMM SPMM backward final fused (CUDA)");
  m.def("backward_fixed64", &spmm_forward, "HCSPMM SPMM backward fixed 64 (CUDA)");
  m.def("backward_fixed64_fused", &spmm_forward_fused, "HCSPMM SPMM backward fixed 64 fused (CUDA)");
  m.def("backward_final_fused_64", &spmm_forward_final_fused, "HCSPMM SPMM backward final fused 64 (CUDA)");
  m.def("backward_GIN_final_fused", &spmm_forward_fused, "HCSPMM SPMM backward for GIN final fused (CUDA)");
  // additions (not in the reference)
  m.def("spmm_strided", &spmm_strided, "out (+)= A @ input; input / out may be row-strided views",
        pybind11::arg("input"), pybind11::arg("nodePointer"), pybind11::arg("edgeList"), pybind11::arg("blockPartition"),
        pybind11::arg("edgeToColumn"), pybind11::arg("edgeToRow"), pybind11::arg("hybrid_type"), pybind11::arg("out"),
        pybind11::arg("accumulate"), pybind11::arg("row_nzr") = pybind11::none(), pybind11::arg("col_nzr") = pybind11::none());
  m.def("spmm_accumulate", &spmm_strided, "alias of spmm_strided",
        pybind11::arg("input"), pybind11::arg("nodePointer"), pybind11::arg("edgeList"), pybind11::arg("blockPartition"),
        pybind11::arg("edgeToColumn"), pybind11::arg("edgeToRow"), pybind11::arg("hybrid_type"), pybind11::arg("out"),
        pybind11::arg("accumulate"), pybind11::arg("row_nzr") = pybind11::none(), pybind11::arg("col_nzr") = pybind11::none());
  m.def("spmm_bf16", &spmm_bf16, "out (+)= A @ input for a bfloat16-stored input (exchange operand)",
        pybind11::arg("input"), pybind11::arg("nodePointer"), pybind11::arg("edgeList"), pybind11::arg("blockPartition"),
        pybind11::arg("edgeToColumn"), pybind11::arg("edgeToRow"), pybind11::arg("hybrid_type"), pybind11::arg("out"),
        pybind11::arg("accumulate"), pybind11::arg("row_nzr") = pybind11::none(), pybind11::arg("col_nzr") = pybind11::none());
  m.def("spmm_segments", &spmm_segments, "out (+)= A @ X with X in segments (peer-mapped operands read in place)",
        pybind11::arg("input"), pybind11::arg("nodePointer"), pybind11::arg("colidx_segments"), pybind11::arg("segment_ptrs"), pybind11::arg("x_rows"),
        pybind11::arg("out"), pybind11::arg("accumulate") = false, pybind11::arg("row_nzr") = pybind11::none(), pybind11::arg("col_nzr") = pybind11::none());
  m.def("f32_to_bf16_into", &f32_to_bf16_into, "out (bfloat16 view) = RNE(input)");
  m.def("gemm_tf32", &gemm_tf32, "a @ b with TF32 tensor-core product (optionally into a row-strided `out`)",
        pybind11::arg("a"), pybind11::arg("b"), pybind11::arg("out") = pybind11::none());
  m.def("set_classifier", &set_classifier, "shipped | intended | b200 | all_cuda | all_tc; returns the previous mode");
  m.def("set_precision", &set_precision, "tf32 | tf32x2 | fp32 | bf16; returns the previous mode");
  m.def("set_dense", &set_dense, "tcgen05 kernels: dense super-window plans in preprocess()/forward*(), Update GEMM");
  m.def("set_row_sort", &set_row_sort, "preprocess() keeps a row-sorted copy of low-degree CSRs for the balanced kernel (default on); returns the previous setting");
  m.def("set_tag_columns", &set_tag_columns, "preprocess() also emits hotness-tagged column ids for the L2 residency hints of wide gathers (default off)");
  m.def("set_bug_compat", &set_bug_compat, "read strided `weights` as raw memory like the reference");
  m.def("set_tuning", &set_tuning, "kernel tuning knob (long_row, slab); returns the previous value");
}
