#!/usr/bin/env python
"""Re-calibration of the core selector on B200 by the reference paper's own recipe (technical report IV-C, p.6-7;
the reference's fitted coefficients are hybrid_all_kernel.cu:261): synthetic row windows with a controlled number of
distinct columns and a controlled sparsity are timed on BOTH paths, and a logistic regression on the selector's two
features separates the regions.

Two granularities:
  window16   16-row windows, 1..130 distinct columns, sparsity 1/16..15/16 (the paper's grid): CUDA-core path vs the
             per-window tensor-core path (mma.sync TF32).  Features as in the reference: (U - 1, density) with
             density = E_w / (ceil(U / 8) * 128); label 1 = CUDA cores faster (score > 0 in the reference's form).
  super128   128-row super-windows (the tcgen05 dense path's unit), 8..1024 distinct columns: CUDA-core path vs the
             tcgen05 dense kernel.  Features (U, density = E / (128 U)); label 1 = tensor cores faster.
Every super-window draws its columns at random from X (no spatial locality, as in the paper's synthetic windows).
Writes the cells and the fitted coefficients to --out (JSON)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hcspmm import capi  # noqa: E402


def synth(rows_per_group: int, groups: int, u: int, density: float, x_rows: int, seed: int, dev):
    """groups x rows_per_group rows; group g uses u distinct pseudo-random columns; cell present with prob density."""
    g = torch.Generator(device=dev).manual_seed(seed)
    base = torch.randint(0, x_rows, (groups, 1), device=dev, generator=g)
    stride = torch.randint(0, x_rows // 2, (groups, 1), device=dev, generator=g) * 2 + 1      # odd: a permutation mod 2^k
    cols = (base + stride * torch.arange(u, device=dev).view(1, u)) % x_rows                 # [groups, u] distinct per group
    cols = torch.sort(cols, dim=1).values
    keep = torch.rand(groups, rows_per_group, u, device=dev, generator=g) < density
    keep[:, :, 0] |= ~keep.any(dim=2)                                                        # no empty rows
    gi, ri, ki = keep.nonzero(as_tuple=True)
    row = gi * rows_per_group + ri
    col = cols[gi, ki]
    n = groups * rows_per_group
    order = torch.argsort(row * x_rows + col)
    row, col = row[order], col[order]
    rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rp[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rp.to(torch.int32), col.to(torch.int32).contiguous()


def timeit(fn, n=8):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def fit(cells, fx, fy, label_key):
    from sklearn.linear_model import LogisticRegression
    x = np.array([[c[fx], c[fy]] for c in cells], dtype=np.float64)
    y = np.array([c[label_key] for c in cells], dtype=np.int64)
    if y.min() == y.max():
        return {"degenerate": True, "label": int(y[0]), "cells": int(y.size),
                "meaning": "one path is faster in EVERY cell of the grid: there is no boundary to fit"}
    m = LogisticRegression(C=1e4, max_iter=10000).fit(x, y)
    acc = float((m.predict(x) == y).mean())
    return {"degenerate": False, "coef_" + fx: float(m.coef_[0][0]), "coef_" + fy: float(m.coef_[0][1]),
            "intercept": float(m.intercept_[0]), "train_accuracy": acc, "cells": int(y.size),
            "positive_fraction": float(y.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_selector_fit.json"))
    ap.add_argument("--dims", default="32,128,256")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dims = [int(d) for d in args.dims.split(",")]
    x_rows = 1 << 18
    res = {"window16": [], "super128": [], "x_rows": x_rows}

    # ---- the paper's grid: 16-row windows ---------------------------------------------------------------------
    us = [1, 2, 4, 8, 16, 24, 32, 48, 64, 96, 130] if not args.quick else [4, 32, 130]
    dens = [1 / 16, 2 / 16, 4 / 16, 6 / 16, 8 / 16, 10 / 16, 12 / 16, 15 / 16] if not args.quick else [1 / 16, 8 / 16, 15 / 16]
    for u in us:
        for dn in dens:
            rp, ci = synth(16, 16384, u, dn, x_rows, 1000 * u + int(dn * 16), dev)
            n = rp.numel() - 1
            pre_c = capi.preprocess(ci, rp, "all_cuda")
            pre_t = capi.preprocess(ci, rp, "all_tc")
            aux_c, aux_t = capi.GraphAux(rp, ci, pre_c[3]), capi.GraphAux(rp, ci, pre_t[3])
            e_w = ci.numel() / (n / 16)
            feat_density = e_w / (((u + 7) // 8) * 128)
            for dim in dims:
                x = torch.randn(x_rows, dim, device=dev)
                t_c = timeit(lambda: capi.spmm_aux(x, rp, ci, *pre_c, aux_c))
                t_t = timeit(lambda: capi.spmm_aux(x, rp, ci, *pre_t, aux_t))
                res["window16"].append({"u": u, "u_minus_1": u - 1, "cell_density": dn, "density": feat_density, "dim": dim,
                                        "entries_per_window": e_w, "cuda_ms": t_c, "tc_ms": t_t, "cuda_faster": int(t_c <= t_t)})
            print(f"window16 u {u} density {dn:.3f}: " +
                  " ".join(f"D{c['dim']} cuda {c['cuda_ms']:.3f} tc {c['tc_ms']:.3f}" for c in res["window16"][-len(dims):]), flush=True)

    # ---- 128-row super-windows: the tcgen05 dense path --------------------------------------------------------
    us = [8, 16, 32, 64, 128, 256, 512, 1024] if not args.quick else [32, 256]
    dens = [1 / 128, 1 / 64, 1 / 32, 1 / 16, 2 / 16, 4 / 16, 8 / 16, 12 / 16, 15 / 16] if not args.quick else [1 / 64, 4 / 16]
    for u in us:
        for dn in dens:
            rp, ci = synth(128, 2048, u, dn, x_rows, 7000 * u + int(dn * 128), dev)
            pre_c = capi.preprocess(ci, rp, "all_cuda")
            pre_t = capi.preprocess(ci, rp, "all_tc")
            plan = capi.DensePlan(rp, ci, pre_t[2], pre_t[3], min_reuse=0.0)
            aux_c, aux_d = capi.GraphAux(rp, ci, pre_c[3]), capi.GraphAux(rp, ci, pre_t[3], plan)
            e_sw = ci.numel() / 2048
            for dim in dims:
                x = torch.randn(x_rows, dim, device=dev)
                t_c = timeit(lambda: capi.spmm_aux(x, rp, ci, *pre_c, aux_c))
                t_d = timeit(lambda: capi.spmm_aux(x, rp, ci, *pre_t, aux_d))
                res["super128"].append({"u": u, "cell_density": dn, "density": e_sw / (128 * u), "reuse": e_sw / u, "dim": dim,
                                        "entries_per_superwindow": e_sw, "cuda_ms": t_c, "dense_ms": t_d,
                                        "dense_faster": int(t_d < t_c)})
            print(f"super128 u {u} density {dn:.4f} reuse {e_sw / u:.1f}: " +
                  " ".join(f"D{c['dim']} cuda {c['cuda_ms']:.3f} dense {c['dense_ms']:.3f}" for c in res["super128"][-len(dims):]), flush=True)

    res["fit_window16"] = {"all_dims": fit(res["window16"], "u_minus_1", "density", "cuda_faster")}
    res["fit_super128"] = {"all_dims": fit(res["super128"], "u", "density", "dense_faster")}
    for dim in dims:
        res["fit_window16"][f"dim{dim}"] = fit([c for c in res["window16"] if c["dim"] == dim], "u_minus_1", "density", "cuda_faster")
        res["fit_super128"][f"dim{dim}"] = fit([c for c in res["super128"] if c["dim"] == dim], "u", "density", "dense_faster")
    # the shipped b200 rule for comparison: dense iff mean column reuse (entries / distinct column) >= 2
    sw = res["super128"]
    res["rule_reuse_ge_2"] = {"agreement": float(np.mean([(c["reuse"] >= 2.0) == bool(c["dense_faster"]) for c in sw])),
                              "best_reuse_threshold": None}
    best = max(((float(np.mean([(c["reuse"] >= th) == bool(c["dense_faster"]) for c in sw])), th)
                for th in (0.5, 1, 1.5, 2, 3, 4, 6, 8, 12, 16, 24, 32)), key=lambda t: t[0])
    res["rule_reuse_ge_2"]["best_reuse_threshold"] = {"threshold": best[1], "agreement": best[0]}
    print(json.dumps({k: res[k] for k in ("fit_window16", "fit_super128", "rule_reuse_ge_2")}, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
