"""ctypes binding of the C ABI (include/hcspmm.h) -- what a non-torch host binds.

torch is used here only to own device memory and to name the current stream; every call goes
through ``libhcspmm.so``'s ``extern "C"`` entry points with raw pointers.  There is no CPU
fallback: if the library is missing, loading raises.
"""
from __future__ import annotations

import ctypes
import os

import torch

from .build import LIB_PATH

BLK_H, BLK_W = 16, 8
CLASSIFIERS = {"shipped": 0, "intended": 1, "b200": 2, "all_cuda": 3, "all_tc": 4, "b200_window": 5}
PRECISIONS = {"tf32": 0, "tf32x2": 1, "fp32": 2, "bf16": 3, "bf16_stored": 4}

EXPORTS = [
    "hcspmm_version", "hcspmm_last_error", "hcspmm_set_tuning",
    "hcspmm_preprocess_workspace_bytes", "hcspmm_preprocess", "hcspmm_spmm", "hcspmm_spmm_gemm",
    "hcspmm_gemm_tf32", "hcspmm_dense_plan_workspace_bytes", "hcspmm_dense_plan_count", "hcspmm_dense_plan_words",
    "hcspmm_dense_plan_fill", "hcspmm_spmm_plan", "hcspmm_debug_umma_error", "hcspmm_loa_workspace_bytes", "hcspmm_loa_reorder", "hcspmm_graph_create", "hcspmm_graph_spmm_host",
    "hcspmm_graph_get_preprocess", "hcspmm_graph_destroy",
    "hcspmm_peer_alloc", "hcspmm_peer_open", "hcspmm_peer_close", "hcspmm_peer_free", "hcspmm_peer_barrier",
    "hcspmm_halo_pull", "hcspmm_halo_pull_rows", "hcspmm_debug_l2_gather",
    "hcspmm_merge_path_count", "hcspmm_merge_path_splits", "hcspmm_spmm_workspace_bytes", "hcspmm_spmm_aux",
    "hcspmm_f32_to_bf16", "hcspmm_spmm_gemm_aux", "hcspmm_halo_push",
    "hcspmm_tag_columns_workspace_bytes", "hcspmm_tag_columns", "hcspmm_row_sort_workspace_bytes", "hcspmm_row_sort",
]

_lib = None
_vp, _i32, _i64, _int, _sz = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int,
                              ctypes.c_size_t)


class HcspmmError(RuntimeError):
    pass


class Aux(ctypes.Structure):
    """hcspmm_aux_t (include/hcspmm.h): per-graph products handed to every aggregation."""
    _fields_ = [("d_splits", ctypes.c_void_p), ("splits_chunk", ctypes.c_int32), ("n_splits", ctypes.c_int32),
                ("n_tc_windows", ctypes.c_int32), ("d_plan", ctypes.c_void_p), ("n_dense", ctypes.c_int32),
                ("plan_full", ctypes.c_int32), ("total_cols", ctypes.c_int64), ("d_workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
                ("d_colidx_tagged", ctypes.c_void_p), ("d_colidx_segments", ctypes.c_void_p),
                ("segment_x", ctypes.c_void_p * 8), ("d_sorted_rowptr", ctypes.c_void_p),
                ("d_sorted_colidx", ctypes.c_void_p), ("d_sorted_row_id", ctypes.c_void_p)]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HcspmmError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                              "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.hcspmm_version.restype = _int
        L.hcspmm_last_error.restype = ctypes.c_char_p
        L.hcspmm_set_tuning.argtypes = [ctypes.c_char_p, _int]
        L.hcspmm_preprocess_workspace_bytes.restype = _sz
        L.hcspmm_preprocess_workspace_bytes.argtypes = [_i32, _i64]
        L.hcspmm_preprocess.argtypes = [_vp, _vp, _i32, _i64, _i32, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
        L.hcspmm_spmm.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _int,
                                  _int, _vp, _i64, _vp]
        L.hcspmm_spmm_gemm.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32,
                                       _int, _vp, _i64, _i32, _vp, _i64, _vp, _i64, _vp]
        L.hcspmm_gemm_tf32.argtypes = [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp]
        L.hcspmm_dense_plan_workspace_bytes.restype = _sz
        L.hcspmm_dense_plan_workspace_bytes.argtypes = [_i32, _i64]
        L.hcspmm_dense_plan_count.argtypes = [_vp, _vp, _vp, _i32, _i64, _int, _vp, _sz, _vp, _vp]
        L.hcspmm_dense_plan_words.restype = _sz
        L.hcspmm_dense_plan_words.argtypes = [_i32, _i32, _i64]
        L.hcspmm_dense_plan_fill.argtypes = [_vp, _vp, _i32, _i64, _vp, _i32, _i64, _vp, _sz, _vp]
        L.hcspmm_spmm_plan.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _int, _int,
                                       _vp, _i64, _vp, _i32, _i64, _vp]
        L.hcspmm_loa_workspace_bytes.restype = _sz
        L.hcspmm_loa_workspace_bytes.argtypes = [_i32, _i64, _i32]
        L.hcspmm_loa_reorder.argtypes = [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _sz, _vp]
        L.hcspmm_graph_create.argtypes = [_vp, _vp, _i32, _i64, _i32, _int, ctypes.POINTER(_vp)]
        L.hcspmm_graph_spmm_host.argtypes = [_vp, _vp, _i32, _int, _vp]
        L.hcspmm_graph_get_preprocess.argtypes = [_vp, _vp, _vp, _vp, _vp]
        L.hcspmm_graph_destroy.argtypes = [_vp]
        L.hcspmm_graph_destroy.restype = None
        L.hcspmm_peer_alloc.argtypes = [_sz, ctypes.POINTER(_vp), _vp]
        L.hcspmm_peer_open.argtypes = [_vp, ctypes.POINTER(_vp)]
        L.hcspmm_peer_close.argtypes = [_vp]
        L.hcspmm_peer_free.argtypes = [_vp]
        L.hcspmm_peer_barrier.argtypes = [_vp, _i32, _i32, _i32, _vp, _vp]
        L.hcspmm_halo_pull.argtypes = [_vp, _i64, _vp, _vp, _i32, ctypes.c_uint64, _i32, _i32, _i32, _i32, _vp, _i64, _vp]
        L.hcspmm_halo_pull_rows.argtypes = [_vp, _i64, _vp, _vp, _vp, _i32, ctypes.c_uint64, _i32, _i32, _i32, _i32, _vp, _i64, _vp]
        L.hcspmm_debug_l2_gather.argtypes = [_vp, _i32, _i32, _i32, _i32, _vp, _vp]
        L.hcspmm_merge_path_count.restype = _sz
        L.hcspmm_merge_path_count.argtypes = [_i32, _i64, _i32]
        L.hcspmm_merge_path_splits.argtypes = [_vp, _i32, _i64, _i32, _vp, _vp]
        L.hcspmm_spmm_workspace_bytes.restype = _sz
        L.hcspmm_spmm_workspace_bytes.argtypes = [_i32, _i64, _i32]
        L.hcspmm_spmm_aux.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _int, _int,
                                      _vp, _i64, ctypes.POINTER(Aux), _vp]
        L.hcspmm_f32_to_bf16.argtypes = [_vp, _i64, _i32, _i32, _vp, _i64, _vp]
        L.hcspmm_tag_columns_workspace_bytes.restype = _sz
        L.hcspmm_tag_columns_workspace_bytes.argtypes = [_i32, _i64]
        L.hcspmm_tag_columns.argtypes = [_vp, _i64, _i32, _vp, _vp, _sz, _vp]
        L.hcspmm_row_sort_workspace_bytes.restype = _sz
        L.hcspmm_row_sort_workspace_bytes.argtypes = [_i32]
        L.hcspmm_row_sort.argtypes = [_vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _vp]
        L.hcspmm_halo_push.argtypes = [_vp, _i64, _vp, _vp, _vp, _i64, _i32, ctypes.c_uint64, _i32, _i32, _i32, _i32, _vp]
        L.hcspmm_spmm_gemm_aux.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _int, _vp, _i64,
                                           _i32, _vp, _i64, _vp, _i64, ctypes.POINTER(Aux), _vp]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise HcspmmError(f"{what} failed (code {rc}): {lib().hcspmm_last_error().decode()}")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def num_windows(n_rows: int) -> int:
    return (n_rows + BLK_H - 1) // BLK_H


def set_tuning(key: str, value: int) -> int:
    return lib().hcspmm_set_tuning(key.encode(), int(value))


def preprocess(colidx: torch.Tensor, rowptr: torch.Tensor, classifier="shipped"):
    """hcspmm_preprocess on device tensors -> (blockPartition, edgeToColumn, edgeToRow, hybrid_type)."""
    assert colidx.is_cuda and rowptr.is_cuda and colidx.dtype == torch.int32 and rowptr.dtype == torch.int32
    n, nnz = rowptr.numel() - 1, colidx.numel()
    w = num_windows(n)
    mode = CLASSIFIERS[classifier] if isinstance(classifier, str) else int(classifier)
    with torch.cuda.device(colidx.device):
        o = dict(dtype=torch.int32, device=colidx.device)
        bp, ht = torch.zeros(w, **o), torch.zeros(w, **o)
        etc, etr = torch.zeros(nnz, **o), torch.zeros(nnz, **o)
        nbytes = lib().hcspmm_preprocess_workspace_bytes(n, nnz)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=colidx.device)
        _check(lib().hcspmm_preprocess(_ptr(colidx), _ptr(rowptr), n, nnz, w, mode, _ptr(bp), _ptr(etc),
                                       _ptr(etr), _ptr(ht), _ptr(ws), nbytes, _stream(colidx)),
               "hcspmm_preprocess")
    return bp, etc, etr, ht


def spmm(x: torch.Tensor, rowptr, colidx, bp=None, etc=None, etr=None, ht=None, precision="tf32",
         out: torch.Tensor | None = None, accumulate: bool = False, n_rows: int | None = None):
    """hcspmm_spmm on device tensors.  x may be a row-strided view (stride(1) == 1)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    n = rowptr.numel() - 1 if n_rows is None else n_rows
    d = x.shape[1]
    if out is None:
        assert not accumulate
        out = torch.empty((n, d), dtype=torch.float32, device=x.device)
    assert out.stride(1) == 1 or d == 1
    prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    with torch.cuda.device(x.device):
        _check(lib().hcspmm_spmm(_ptr(x), x.stride(0), x.shape[0], _ptr(rowptr), _ptr(colidx), _ptr(bp),
                                 _ptr(etc), _ptr(etr), _ptr(ht), n, colidx.numel(), d, prec,
                                 1 if accumulate else 0, _ptr(out), out.stride(0), _stream(x)),
               "hcspmm_spmm")
    return out


SPLITS_CHUNK = 4096      # HCSPMM_SPLITS_CHUNK


def tag_column_ids(colidx: torch.Tensor, n_cols: int) -> torch.Tensor:
    """hcspmm_tag_columns: the column ids with the hotness class of their column (rank by reference count) in bits 29..31."""
    nnz = colidx.numel()
    out = torch.empty(nnz, dtype=torch.int32, device=colidx.device)
    with torch.cuda.device(colidx.device):
        nbytes = lib().hcspmm_tag_columns_workspace_bytes(n_cols, nnz)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=colidx.device)
        _check(lib().hcspmm_tag_columns(_ptr(colidx), nnz, n_cols, _ptr(out), _ptr(ws), nbytes, _stream(colidx)),
               "hcspmm_tag_columns")
    return out


def want_row_sort(n_rows: int, nnz: int) -> bool:
    """The rule preprocess() applies: graphs the balanced kernel serves (mean row >= 8 entries) whose rows are short
    enough (mean < 64) that an item holds hundreds of them -- there grouping rows of similar length pays (products
    shape: -15 %); on high-degree graphs the warp runs already balance an item."""
    return 8 * n_rows <= nnz < 64 * n_rows


def row_sort_csr(rowptr: torch.Tensor, colidx: torch.Tensor):
    """hcspmm_row_sort -> (row_id [n], sorted_rowptr [n + 1], sorted_colidx [nnz])."""
    n, nnz = rowptr.numel() - 1, colidx.numel()
    dev = rowptr.device
    row_id = torch.empty(n, dtype=torch.int32, device=dev)
    rp_s = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ci_s = torch.empty(nnz, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = lib().hcspmm_row_sort_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _check(lib().hcspmm_row_sort(_ptr(rowptr), _ptr(colidx), n, nnz, _ptr(row_id), _ptr(rp_s), _ptr(ci_s), _ptr(ws), nbytes,
                                     _stream(rowptr)), "hcspmm_row_sort")
    return row_id, rp_s, ci_s


class GraphAux:
    """The per-graph products of hcspmm_aux_t for a device CSR: merge-path split points (computed once), the count of
    windows labelled 1 and a reusable workspace -- what HCSPMM.preprocess() packs into its two opaque tensors."""

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, ht: torch.Tensor | None = None, plan=None,
                 tag_columns: bool = False, n_cols: int | None = None, row_sort: bool | None = None):
        """row_sort: keep a row-sorted copy of the CSR for the balanced kernel (hcspmm_row_sort: rows grouped by the power
        of two of their length; invisible to the caller).  None = the library's rule: low-degree graphs served by the
        balanced kernel (8 <= mean row < 64 entries)."""
        n, nnz = rowptr.numel() - 1, colidx.numel()
        self.n, self.nnz = n, nnz
        self.tagged = tag_column_ids(colidx, n_cols if n_cols is not None else n) if (tag_columns and nnz > 0) else None
        if row_sort is None:
            row_sort = want_row_sort(n, nnz)
        self.sorted = row_sort_csr(rowptr, colidx) if (row_sort and n > 0 and self.tagged is None) else None
        with torch.cuda.device(rowptr.device):
            cnt = lib().hcspmm_merge_path_count(n, nnz, SPLITS_CHUNK)
            self.splits = torch.empty(cnt, dtype=torch.int32, device=rowptr.device)
            rp_for_splits = self.sorted[1] if self.sorted is not None else rowptr
            _check(lib().hcspmm_merge_path_splits(_ptr(rp_for_splits), n, nnz, SPLITS_CHUNK, _ptr(self.splits), _stream(rowptr)),
                   "hcspmm_merge_path_splits")
        self.n_tc = int((ht == 1).sum()) if ht is not None else -1
        self.plan = plan
        self.plan_full = 0
        if plan is not None and plan.n_dense > 0:        # does the plan cover every super-window that has entries?
            edges = torch.arange(0, n + 128, 128, device=rowptr.device).clamp_(max=n)
            self.plan_full = int(int((rowptr[edges[1:]] > rowptr[edges[:-1]]).sum()) == plan.n_dense)
        self._ws = {}

    def struct(self, dim: int, device) -> Aux:
        if dim not in self._ws:
            nbytes = lib().hcspmm_spmm_workspace_bytes(self.n, self.nnz, dim)
            self._ws[dim] = torch.empty(nbytes, dtype=torch.uint8, device=device)
        ws = self._ws[dim]
        a = Aux()
        a.d_splits, a.splits_chunk, a.n_splits = self.splits.data_ptr(), SPLITS_CHUNK, self.splits.numel() - 1
        a.n_tc_windows = self.n_tc
        if self.plan is not None:
            a.d_plan, a.n_dense, a.total_cols = self.plan.plan.data_ptr(), self.plan.n_dense, self.plan.total_cols
            a.plan_full = self.plan_full
        a.d_workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        if self.tagged is not None:
            a.d_colidx_tagged = self.tagged.data_ptr()
        if self.sorted is not None:
            a.d_sorted_row_id, a.d_sorted_rowptr, a.d_sorted_colidx = (t.data_ptr() for t in self.sorted)
        return a


def spmm_aux(x: torch.Tensor, rowptr, colidx, bp, etc, etr, ht, aux: GraphAux, precision="tf32",
             out: torch.Tensor | None = None, accumulate: bool = False):
    """hcspmm_spmm_aux on device tensors (x: FP32, or bfloat16 with precision "bf16_stored")."""
    n, d = rowptr.numel() - 1, x.shape[1]
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=x.device)
    prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    a = aux.struct(d, x.device)
    with torch.cuda.device(x.device):
        _check(lib().hcspmm_spmm_aux(_ptr(x), x.stride(0), x.shape[0], _ptr(rowptr), _ptr(colidx), _ptr(bp), _ptr(etc),
                                     _ptr(etr), _ptr(ht), n, colidx.numel(), d, prec, 1 if accumulate else 0, _ptr(out),
                                     out.stride(0), ctypes.byref(a), _stream(x)),
               "hcspmm_spmm_aux")
    return out


def spmm_gemm_aux(x, rowptr, colidx, bp, etc, etr, ht, w: torch.Tensor, aux: GraphAux, precision="tf32"):
    """hcspmm_spmm_gemm_aux -> (out [n, hidden], z [n, dim]); one fused kernel when the dense plan covers the graph."""
    assert x.is_cuda and x.is_contiguous() and w.is_cuda and w.is_contiguous()
    n, d, h = rowptr.numel() - 1, x.shape[1], w.shape[1]
    out = torch.empty((n, h), dtype=torch.float32, device=x.device)
    z = torch.empty((n, d), dtype=torch.float32, device=x.device)
    prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    a = aux.struct(d, x.device)
    with torch.cuda.device(x.device):
        _check(lib().hcspmm_spmm_gemm_aux(_ptr(x), d, x.shape[0], _ptr(rowptr), _ptr(colidx), _ptr(bp), _ptr(etc),
                                          _ptr(etr), _ptr(ht), n, colidx.numel(), d, prec, _ptr(w), h, h,
                                          _ptr(out), h, _ptr(z), d, ctypes.byref(a), _stream(x)),
               "hcspmm_spmm_gemm_aux")
    return out, z


def f32_to_bf16(x: torch.Tensor, out: torch.Tensor | None = None):
    """hcspmm_f32_to_bf16: round-to-nearest-even copy of FP32 rows into a bfloat16 tensor (any row pitch)."""
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().hcspmm_f32_to_bf16(_ptr(x), x.stride(0), x.shape[0], x.shape[1], _ptr(out), out.stride(0),
                                        _stream(x)), "hcspmm_f32_to_bf16")
    return out


class DensePlan:
    """hcspmm_dense_plan_*: the tcgen05 dense super-window plan of a preprocessed graph."""

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, etr: torch.Tensor, ht: torch.Tensor,
                 min_reuse: float = 2.0):
        n, nnz = rowptr.numel() - 1, colidx.numel()
        self.n_rows = n
        with torch.cuda.device(rowptr.device):
            nbytes = lib().hcspmm_dense_plan_workspace_bytes(n, nnz)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=rowptr.device)
            counts = (ctypes.c_int32 * 2)()
            _check(lib().hcspmm_dense_plan_count(_ptr(colidx), _ptr(rowptr), _ptr(ht), n, nnz, int(round(min_reuse * 2)),
                                                 _ptr(ws), nbytes, ctypes.addressof(counts), _stream(rowptr)),
                   "hcspmm_dense_plan_count")
            self.n_dense, self.total_cols = int(counts[0]), int(counts[1])
            words = lib().hcspmm_dense_plan_words(n, self.n_dense, self.total_cols)
            self.plan = torch.zeros(words, dtype=torch.int32, device=rowptr.device)
            _check(lib().hcspmm_dense_plan_fill(_ptr(colidx), _ptr(etr), n, nnz, _ptr(ws), self.n_dense, self.total_cols,
                                                _ptr(self.plan), words, _stream(rowptr)), "hcspmm_dense_plan_fill")


def spmm_plan(x, rowptr, colidx, bp, etc, etr, ht, plan: DensePlan, precision="tf32", out=None, accumulate=False):
    """hcspmm_spmm_plan on device tensors."""
    n, d = rowptr.numel() - 1, x.shape[1]
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=x.device)
    prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    with torch.cuda.device(x.device):
        _check(lib().hcspmm_spmm_plan(_ptr(x), x.stride(0), x.shape[0], _ptr(rowptr), _ptr(colidx), _ptr(bp), _ptr(etc),
                                      _ptr(etr), _ptr(ht), n, colidx.numel(), d, prec, 1 if accumulate else 0, _ptr(out),
                                      out.stride(0), _ptr(plan.plan), plan.n_dense, plan.total_cols, _stream(x)),
               "hcspmm_spmm_plan")
    return out


def spmm_gemm(x, rowptr, colidx, bp, etc, etr, ht, w: torch.Tensor, precision="tf32"):
    """hcspmm_spmm_gemm -> (out [n, hidden], z [n, dim])."""
    assert x.is_cuda and x.is_contiguous() and w.is_cuda and w.is_contiguous()
    n, d, h = rowptr.numel() - 1, x.shape[1], w.shape[1]
    out = torch.empty((n, h), dtype=torch.float32, device=x.device)
    z = torch.empty((n, d), dtype=torch.float32, device=x.device)
    prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    with torch.cuda.device(x.device):
        _check(lib().hcspmm_spmm_gemm(_ptr(x), d, x.shape[0], _ptr(rowptr), _ptr(colidx), _ptr(bp), _ptr(etc),
                                      _ptr(etr), _ptr(ht), n, colidx.numel(), d, prec, _ptr(w), h, h,
                                      _ptr(out), h, _ptr(z), d, _stream(x)),
               "hcspmm_spmm_gemm")
    return out, z


def gemm_tf32(a: torch.Tensor, b: torch.Tensor):
    assert a.is_cuda and b.is_cuda and a.stride(1) == 1 and b.stride(1) == 1
    out = torch.empty((a.shape[0], b.shape[1]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _check(lib().hcspmm_gemm_tf32(_ptr(a), a.stride(0), _ptr(b), b.stride(0), a.shape[0], a.shape[1],
                                      b.shape[1], _ptr(out), out.stride(0), _stream(a)),
               "hcspmm_gemm_tf32")
    return out


def l2_gather_bandwidth(device, row_floats: int = 256, megabytes: int = 32, iters: int = 512, reps: int = 5) -> float:
    """Measured L2 -> SM gather bandwidth in GB/s (hcspmm_debug_l2_gather): random rows of an L2-resident
    [rows, row_floats] buffer, the SpMM's 256-bit evict_last loads, best of `reps` timed launches."""
    rows = megabytes * (1 << 20) // (row_floats * 4)
    with torch.cuda.device(device):
        buf = torch.zeros(rows, row_floats, device=device)
        sink = torch.zeros(1, device=device)
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        ctas = sms * 2 * 4
        iters = (iters + 7) // 8 * 8
        st = torch.cuda.current_stream(device).cuda_stream

        def go():
            _check(lib().hcspmm_debug_l2_gather(buf.data_ptr(), rows, row_floats, iters, ctas, sink.data_ptr(), st),
                   "hcspmm_debug_l2_gather")
        go(); go()
        best = float("inf")
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); go(); b.record()
            torch.cuda.synchronize(device)
            best = min(best, a.elapsed_time(b))
    return ctas * 8 * iters * row_floats * 4 / (best * 1e-3) / 1e9


def csc_of(rowptr: torch.Tensor, colidx: torch.Tensor):
    """CSC of a device CSR with ascending row lists -- what LOI.cpp's main builds (:826-841)."""
    n = rowptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device=rowptr.device), (rowptr[1:] - rowptr[:-1]).long())
    key = torch.sort(colidx.long() * n + rows).values
    cnt = torch.bincount(torch.div(key, n, rounding_mode="floor"), minlength=n)
    rp_in = torch.zeros(n + 1, dtype=torch.int64, device=rowptr.device)
    rp_in[1:] = torch.cumsum(cnt, 0)
    return rp_in.to(torch.int32), (key % n).to(torch.int32)


def loa_reorder(rowptr: torch.Tensor, colidx: torch.Tensor):
    """hcspmm_loa_reorder on a device CSR -> (perm int32[n], block_sizes int32[n_blocks], n_full)."""
    assert rowptr.is_cuda and colidx.is_cuda and rowptr.dtype == torch.int32 and colidx.dtype == torch.int32
    n, nnz = rowptr.numel() - 1, colidx.numel()
    if n == 0:
        return torch.zeros(0, dtype=torch.int32, device=rowptr.device), torch.zeros(0, dtype=torch.int32), 0
    rp_in, ci_in = csc_of(rowptr, colidx)
    maxdeg = int((rowptr[1:] - rowptr[:-1]).max())
    with torch.cuda.device(rowptr.device):
        o = dict(dtype=torch.int32, device=rowptr.device)
        perm, bstart, counts = torch.empty(n, **o), torch.zeros(n + 1, **o), torch.zeros(2, **o)
        nbytes = lib().hcspmm_loa_workspace_bytes(n, nnz, maxdeg)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=rowptr.device)
        _check(lib().hcspmm_loa_reorder(_ptr(rowptr), _ptr(colidx), _ptr(rp_in), _ptr(ci_in), n, nnz, maxdeg,
                                        _ptr(perm), _ptr(bstart), _ptr(counts), _ptr(ws), nbytes, _stream(rowptr)),
               "hcspmm_loa_reorder")
        nb, nf = (int(v) for v in counts.cpu())
        sizes = (bstart[1:nb + 1] - bstart[:nb]).cpu()
    return perm, sizes, nf


def relabel(rowptr: torch.Tensor, colidx: torch.Tensor, perm: torch.Tensor):
    """Apply an LOA permutation (perm[new] = old) symmetrically to rows and columns and re-canonicalise
    the CSR -- the 'Reorder G using NRW' step the reference never shipped."""
    from . import graphs
    n = rowptr.numel() - 1
    inv = torch.empty(n, dtype=torch.int64, device=rowptr.device)
    inv[perm.long()] = torch.arange(n, device=rowptr.device)
    rows = torch.repeat_interleave(torch.arange(n, device=rowptr.device), (rowptr[1:] - rowptr[:-1]).long())
    return graphs.csr_from_pairs(inv[rows], inv[colidx.long()], n, symmetrize=False)


class HostGraph:
    """hcspmm_graph_*: the host-buffer path (CSR and X/Y live in host memory; the library owns the
    device copies and runs H2D -> kernel -> D2H inside each call)."""

    def __init__(self, rowptr_cpu: torch.Tensor, colidx_cpu: torch.Tensor, x_rows: int | None = None,
                 classifier="shipped"):
        assert not rowptr_cpu.is_cuda and rowptr_cpu.dtype == torch.int32
        self.n_rows = rowptr_cpu.numel() - 1
        self.nnz = colidx_cpu.numel()
        self.x_rows = self.n_rows if x_rows is None else x_rows
        self._h = _vp()
        mode = CLASSIFIERS[classifier] if isinstance(classifier, str) else int(classifier)
        _check(lib().hcspmm_graph_create(rowptr_cpu.contiguous().data_ptr(), colidx_cpu.contiguous().data_ptr(),
                                         self.n_rows, self.nnz, self.x_rows, mode, ctypes.byref(self._h)),
               "hcspmm_graph_create")

    def spmm(self, x_cpu: torch.Tensor, out_cpu: torch.Tensor | None = None, precision="tf32"):
        assert not x_cpu.is_cuda and x_cpu.is_contiguous() and x_cpu.shape[0] == self.x_rows
        d = x_cpu.shape[1]
        if out_cpu is None:
            out_cpu = torch.empty((self.n_rows, d), dtype=torch.float32)
        prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        _check(lib().hcspmm_graph_spmm_host(self._h, x_cpu.data_ptr(), d, prec, out_cpu.data_ptr()),
               "hcspmm_graph_spmm_host")
        return out_cpu

    def preprocess_arrays(self):
        w = num_windows(self.n_rows)
        bp, ht = torch.zeros(w, dtype=torch.int32), torch.zeros(w, dtype=torch.int32)
        etc, etr = torch.zeros(self.nnz, dtype=torch.int32), torch.zeros(self.nnz, dtype=torch.int32)
        _check(lib().hcspmm_graph_get_preprocess(self._h, bp.data_ptr(), etc.data_ptr(), etr.data_ptr(),
                                                 ht.data_ptr()), "hcspmm_graph_get_preprocess")
        return bp, etc, etr, ht

    def close(self):
        if self._h:
            lib().hcspmm_graph_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
