"""Shared test helpers: small seeded graphs (numpy) and error metrics."""
import numpy as np
import torch

from hcspmm import graphs


def small_graphs():
    """name -> (rowptr, colidx) int32 numpy.  Covers: regular low degree (reference envelope),
    banded (few distinct columns per window), power law with hubs, N % 16 != 0, empty rows,
    empty windows, a fully empty graph and duplicate-heavy dense windows."""
    out = {}
    for name, (rp, ci) in {
        "ring3_256": graphs.ring_matching(256, seed=0),
        "band2_320": graphs.banded(320, 2),
        "rmat_1000": graphs.rmat(1000, 16000, seed=11),          # N % 16 = 8, hubs
        "rmat_hub_4096": graphs.rmat(4096, 200000, seed=12, a=0.7, b=0.12, c=0.12),
        "sbm_1024": graphs.sbm_dense_windows(1024, community=128, p_in=0.5, extra=4, seed=3),
        "uniform_777": graphs.uniform_random(777, 6, seed=5),
    }.items():
        out[name] = (rp.numpy().astype(np.int32), ci.numpy().astype(np.int32))
    # empty rows and empty windows: rows 40..120 have no edges
    rp, ci = out["uniform_777"]
    deg = np.diff(rp).copy()
    keep = np.ones(ci.size, bool)
    for r in range(40, 121):
        keep[rp[r]:rp[r + 1]] = False
        deg[r] = 0
    rp2 = np.zeros_like(rp)
    rp2[1:] = np.cumsum(deg)
    out["holes_777"] = (rp2.astype(np.int32), ci[keep].astype(np.int32))
    out["empty_48"] = (np.zeros(49, np.int32), np.zeros(0, np.int32))
    out["single_row"] = (np.array([0, 1], np.int32), np.array([0], np.int32))
    # > 4096 edges per window: the bitmap path of the GPU preprocessing
    rp, ci = graphs.uniform_random(2048, 400, seed=21)
    out["dense_2048"] = (rp.numpy().astype(np.int32), ci.numpy().astype(np.int32))
    # rectangular shard: 72 rows, global column ids up to 100000, ~700 edges per row
    # (window column span >> bitmap chunk => several chunks per window)
    rng = np.random.default_rng(31)
    rows = []
    for r in range(72):
        k = 0 if r in (5, 6) else int(rng.integers(500, 900))
        rows.append(np.sort(rng.choice(100000, size=k, replace=False)).astype(np.int32))
    rp = np.zeros(73, np.int32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    out["rect_72x100000"] = (rp, np.concatenate(rows).astype(np.int32))
    return out


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def rel_max(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max() if b.size else 1.0
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0)) if a.size else 0.0


def torch_sparse_ref(rowptr, colidx, x, n_rows=None):
    """FP32 torch.sparse.mm on CPU -- the north star's stated oracle."""
    n = rowptr.size - 1 if n_rows is None else n_rows
    a = torch.sparse_csr_tensor(torch.from_numpy(rowptr.astype(np.int64)),
                                torch.from_numpy(colidx.astype(np.int64)),
                                torch.ones(colidx.size, dtype=torch.float32), size=(n, x.shape[0]))
    return torch.sparse.mm(a, torch.from_numpy(np.ascontiguousarray(x))).numpy()
