#!/usr/bin/env python
"""Time the single-GPU SpMM kernel over tuning-knob settings on one shape, in ONE process (graph built
once).  usage: sweep_kernel.py --shape reddit --dim 256 --rows-frac 1.0 --set balance=0 --set balance=1,chunk=2048 ...
--rows-frac f keeps the first f of the rows (a 1/P row shard as a multi-GPU rank sees it)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hc-spmm_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="reddit")
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--rows-frac", type=float, default=1.0)
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--classifier", default="shipped")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--set", action="append", default=[])
    args = ap.parse_args()
    import HCSPMM
    from hcspmm import graphs, partition
    dev = torch.device("cuda", 0)
    rp, ci, info = graphs.named(args.shape, device=dev)
    n = info["n"]
    dim = args.dim or info["dim"]
    if args.rows_frac < 1.0:
        world = round(1.0 / args.rows_frac)
        cuts = partition.window_cuts(rp, world)
        rp, ci = partition.local_shard(rp, ci, cuts[0], cuts[1])
    n_l, nnz = rp.numel() - 1, ci.numel()
    HCSPMM.set_classifier(args.classifier)
    HCSPMM.set_precision(args.precision)
    pre = HCSPMM.preprocess(ci, rp, n_l, nnz, (n_l + 15) // 16)
    x = torch.randn(n, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
    ref = None
    for spec in args.set or [""]:
        kv = dict(s.split("=") for s in spec.split(",") if s)
        old = {k: HCSPMM.set_tuning(k, int(v)) for k, v in kv.items()}
        for _ in range(3):
            y = HCSPMM.forward(x, rp, ci, *pre)[0]
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            a.record()
            y = HCSPMM.forward(x, rp, ci, *pre)[0]
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        if ref is None:
            ref = y.clone()
        err = float((y - ref).norm() / ref.norm())
        for k, v in old.items():
            HCSPMM.set_tuning(k, v)
        print(json.dumps({"shape": args.shape, "dim": dim, "rows": n_l, "nnz": nnz, "set": spec, "precision": args.precision,
                          "ms_median": ms[len(ms) // 2], "ms_min": ms[0], "gflops": 2.0 * nnz * dim / ms[len(ms) // 2] / 1e6,
                          "rel_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()
