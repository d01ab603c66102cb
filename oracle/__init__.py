"""CPU oracle for the HC-SpMM hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``hc-spmm_b200/``) never does; it fails loudly when its CUDA library is missing.

Thin ctypes bindings over ``oracle/_build/liboracle.so`` (plain-C restatement of the
reference, ``oracle/hcspmm_oracle.c``) and, when it was built in the development
container, ``oracle/_ref/libloi_ref.so`` (the unmodified reference ``LOI.cpp``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_LOI_REF = os.path.join(_HERE, "_ref", "libloi_ref.so")

BLK_H = 16
BLK_W = 8
MODE_SHIPPED = 0   # hybrid_all_kernel.cu:262
MODE_INTENDED = 1  # hybrid_all_kernel.cu:261

_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    """Compile the C restatement (and the reference LOA library if the reference is present)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "hcspmm_oracle.c")):
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], env=env,
                              stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.hcspmm_oracle_score.restype = ctypes.c_double
        _lib.hcspmm_oracle_score.argtypes = [ctypes.c_int32, ctypes.c_uint32, ctypes.c_int32]
        _lib.hcspmm_oracle_label.restype = ctypes.c_int32
        _lib.hcspmm_oracle_label.argtypes = [ctypes.c_int32, ctypes.c_uint32, ctypes.c_int32,
                                             ctypes.c_int]
        _lib.hcspmm_oracle_tf32.restype = ctypes.c_float
        _lib.hcspmm_oracle_tf32.argtypes = [ctypes.c_float]
    return _lib


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _p(a: np.ndarray, t=_i32p):
    return a.ctypes.data_as(t)


def num_windows(n_rows: int) -> int:
    return (n_rows + BLK_H - 1) // BLK_H


def preprocess(colidx, rowptr, mode: int = MODE_SHIPPED):
    """Restates HCSPMM.preprocess (hybrid_all_kernel.cu:339-408).

    Returns (blockPartition[W], edgeToColumn[nnz], edgeToRow[nnz], hybrid_type[W]) int32.
    """
    colidx, rowptr = _i32(colidx), _i32(rowptr)
    n = rowptr.size - 1
    nnz = int(rowptr[-1])
    w = num_windows(n)
    bp = np.zeros(w, np.int32)
    etc = np.zeros(max(nnz, 1), np.int32)[:nnz]
    etr = np.zeros(max(nnz, 1), np.int32)[:nnz]
    ht = np.zeros(w, np.int32)
    rc = lib().hcspmm_oracle_preprocess(_p(colidx), _p(rowptr), ctypes.c_int32(n),
                                        ctypes.c_int64(nnz), ctypes.c_int32(w),
                                        ctypes.c_int(mode), _p(bp), _p(etc), _p(etr), _p(ht))
    assert rc == 0
    return bp, etc, etr, ht


def spmm(rowptr, colidx, x, hybrid_type=None, precision: int = 0, n_rows=None, y_init=None):
    """Y = A @ X for binary CSR A (hybrid_all_kernel.cu:982-990, 1099-1111, 1371-1382).

    precision 0: FP32 on CUDA-core windows, TF32-rounded X on tensor-core windows;
    1: FP32 everywhere; 2: BF16-rounded X everywhere.  y_init: accumulate into a copy.
    """
    colidx, rowptr = _i32(colidx), _i32(rowptr)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = rowptr.size - 1 if n_rows is None else n_rows
    d = x.shape[1]
    beta = 0
    if y_init is not None:
        y = np.array(y_init, dtype=np.float32, order="C", copy=True)
        beta = 1
    else:
        y = np.empty((n, d), np.float32)
    ht = None if hybrid_type is None else _i32(hybrid_type)
    rc = lib().hcspmm_oracle_spmm(_p(rowptr), _p(colidx), _p(ht) if ht is not None else None,
                                  ctypes.c_int32(n), ctypes.c_int32(x.shape[0]),
                                  ctypes.c_int32(d), _p(x, _f32p), ctypes.c_int64(d),
                                  _p(y, _f32p), ctypes.c_int64(d), ctypes.c_int(precision),
                                  ctypes.c_int(beta))
    assert rc == 0
    return y


def gemm(z, w, tf32: bool = False):
    """out = Z @ W (hybrid_all_kernel.cu:1809-1837 when tf32)."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.empty((z.shape[0], w.shape[1]), np.float32)
    rc = lib().hcspmm_oracle_gemm(_p(z, _f32p), ctypes.c_int64(z.shape[1]), _p(w, _f32p),
                                  ctypes.c_int64(w.shape[1]), ctypes.c_int32(z.shape[0]),
                                  ctypes.c_int32(z.shape[1]), ctypes.c_int32(w.shape[1]),
                                  _p(out, _f32p), ctypes.c_int64(w.shape[1]),
                                  ctypes.c_int(1 if tf32 else 0))
    assert rc == 0
    return out


def tf32_round(x) -> np.ndarray:
    """cvt.rna.tf32.f32 on an array (numpy restatement of hcspmm_oracle_tf32)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).copy()
    fin = (u & 0x7F800000) != 0x7F800000
    u[fin] = (u[fin] + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def csc(rowptr, colidx):
    rowptr, colidx = _i32(rowptr), _i32(colidx)
    n = rowptr.size - 1
    rp_in = np.zeros(n + 1, np.int32)
    ci_in = np.zeros(max(int(rowptr[-1]), 1), np.int32)
    lib().hcspmm_oracle_csc(_p(rowptr), _p(colidx), ctypes.c_int32(n), _p(rp_in), _p(ci_in))
    return rp_in, ci_in[: int(rowptr[-1])]


def _loa_call(fn, rowptr, colidx, with_csc: bool):
    rowptr, colidx = _i32(rowptr), _i32(colidx)
    n = rowptr.size - 1
    perm = np.full(max(n, 1), -1, np.int32)
    sizes = np.zeros(max(n, 1), np.int32)
    nb, nf = ctypes.c_int32(0), ctypes.c_int32(0)
    if with_csc:
        rp_in, ci_in = csc(rowptr, colidx)
        ci_in = np.ascontiguousarray(np.concatenate([ci_in, np.zeros(1, np.int32)]))
        rc = fn(_p(rowptr), _p(colidx), _p(rp_in), _p(ci_in), ctypes.c_int32(n), _p(perm),
                _p(sizes), ctypes.byref(nb), ctypes.byref(nf))
    else:
        rc = fn(_p(rowptr), _p(colidx), ctypes.c_int32(n), _p(perm), _p(sizes),
                ctypes.byref(nb), ctypes.byref(nf))
    assert rc == 0, rc
    return perm[:n], sizes[: nb.value].copy(), nf.value


def loa(rowptr, colidx):
    """LOA permutation, restating LOI.cpp:660-805 + 873-891.  -> (perm, block_sizes, n_full)."""
    return _loa_call(lib().hcspmm_oracle_loa, rowptr, colidx, True)


def have_loi_ref() -> bool:
    return os.path.exists(_LOI_REF)


_loi = None


def loa_reference(rowptr, colidx):
    """The unmodified reference LOI.cpp (oracle/_ref/libloi_ref.so)."""
    global _loi
    if _loi is None:
        _loi = ctypes.CDLL(_LOI_REF)
    return _loa_call(_loi.loi_ref_reorder, rowptr, colidx, False)
