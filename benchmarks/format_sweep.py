#!/usr/bin/env python
"""Format-selection sweep (BASELINE.json configs[4]): power-law vs banded synthetic graphs,
every core selector, several feature widths.  Prints one JSON line per (graph, dim) with the label
histogram and the SpMM time under
  all_cuda | all_tc (mma.sync per window) | intended (reference coefficients, hybrid_all_kernel.cu:261) |
  b200 (re-fit) | b200+dense (tcgen05 super-windows)
  python benchmarks/format_sweep.py [--nnz 1000000 10000000] [--dims 32 64 128 256] [--quick]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nnz", type=int, nargs="+", default=[1_000_000, 10_000_000])
    ap.add_argument("--deg", type=int, nargs="+", default=[16, 128])
    ap.add_argument("--dims", type=int, nargs="+", default=[32, 64, 128, 256, 512])
    ap.add_argument("--bands", type=int, nargs="+", default=[32, 256, 2048])
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    if args.quick:
        args.nnz, args.deg, args.dims, args.bands = [1_000_000], [16], [32, 256], [32, 2048]
    import HCSPMM
    from hcspmm import graphs
    dev = torch.device("cuda", 0)
    modes = [("all_cuda", False), ("all_tc", False), ("intended", False), ("b200", False), ("b200", True)]
    for nnz in args.nnz:
        for deg in args.deg:
            n = max(1024, nnz // deg // 16 * 16)
            specs = [("rmat", lambda: graphs.rmat(n, nnz, seed=5, device=dev))]
            specs += [(f"band{b}", (lambda b=b: graphs.banded_random(n, deg, b, seed=5, device=dev))) for b in args.bands]
            for gname, make in specs:
                rp, ci = make()
                nn, ne = rp.numel() - 1, ci.numel()
                for dim in args.dims:
                    x = torch.randn(nn, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
                    row = {"graph": gname, "nodes": nn, "stored_entries": ne, "avg_degree": ne / nn, "dim": dim}
                    ref = None
                    for mode, dense in modes:
                        HCSPMM.set_dense(dense)
                        HCSPMM.set_classifier(mode)
                        pre = HCSPMM.preprocess(ci, rp, nn, ne, (nn + 15) // 16)
                        key = mode + ("+dense" if dense else "")
                        ms = timeit(lambda: HCSPMM.forward(x, rp, ci, *pre))
                        y = HCSPMM.forward(x, rp, ci, *pre)[0]
                        if ref is None:
                            ref = y
                        err = float((y - ref).norm() / ref.norm())
                        assert err <= 1e-3, (key, err)
                        row[key] = {"ms": round(ms, 4), "gflops": round(2.0 * ne * dim / ms / 1e6, 1),
                                    "tc_windows": int((pre[3] != 0).sum()),
                                    "dense_groups": int(pre[4][1]) if pre[4].device.type == "cpu" else 0,
                                    "rel_err_vs_all_cuda": err}
                    row["windows"] = (nn + 15) // 16
                    print(json.dumps(row), flush=True)
    HCSPMM.set_dense(False)
    HCSPMM.set_classifier("shipped")


if __name__ == "__main__":
    main()
