#!/usr/bin/env python
"""Format-selection sweep (BASELINE.json configs[4]): power-law vs banded synthetic graphs,
every core selector, several feature widths.  Prints one JSON line per (graph, dim) with the label
histogram and the SpMM time under
  all_cuda | all_tc (mma.sync per window) | intended (reference coefficients, hybrid_all_kernel.cu:261) |
  b200 (re-fit) | b200+dense (tcgen05 super-windows)
  python benchmarks/format_sweep.py [--nnz 1000000 10000000] [--dims 32 64 128 256] [--quick]
  torchrun --nproc-per-node 8 benchmarks/format_sweep.py --nnz 100000000 500000000 ...   (configs[4] is quoted at 8 GPUs:
      the graph is row-window partitioned, every selector preprocesses its shard, the time is one aggregation --
      halo exchange + local SpMM -- max over ranks)
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
    ap.add_argument("--skip-all-tc", action="store_true", help="the per-window tensor-core path is up to 14 x slower on hub windows")
    args = ap.parse_args()
    if args.quick:
        args.nnz, args.deg, args.dims, args.bands = [1_000_000], [16], [32, 256], [32, 2048]
    import HCSPMM
    from hcspmm import graphs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        import torch.distributed as dist
        from hcspmm import dist as hd
        dist.init_process_group("nccl", device_id=dev)
    modes = [("all_cuda", False), ("all_tc", False), ("intended", False), ("b200_window", False), ("b200", False), ("b200", True)]
    if args.skip_all_tc:
        modes = [m for m in modes if m[0] != "all_tc"]
    for nnz in args.nnz:
        for deg in args.deg:
            n = max(1024, nnz // deg // 16 * 16)
            specs = [("rmat", lambda: graphs.rmat(n, nnz, seed=5, device=dev))]
            specs += [(f"band{b}", (lambda b=b: graphs.banded_random(n, deg, b, seed=5, device=dev))) for b in args.bands]
            for gname, make in specs:
                rp, ci = make()
                nn, ne = rp.numel() - 1, ci.numel()
                rows = {dim: {"graph": gname, "nodes": nn, "stored_entries": ne, "avg_degree": ne / nn, "dim": dim,
                              "n_gpus": world, "windows": (nn + 15) // 16} for dim in args.dims}
                refs = {}
                for mode, dense in modes:            # one preprocess (and, at N > 1, one exchange set-up) per selector
                    HCSPMM.set_dense(dense)
                    HCSPMM.set_classifier(mode)
                    key = mode + ("+dense" if dense else "")
                    if world == 1:
                        pre = HCSPMM.preprocess(ci, rp, nn, ne, (nn + 15) // 16)
                    else:
                        sg = hd.ShardedGraph(rp, ci, schedule="auto")
                        pre = sg.pre
                    tcw, dg = int((pre[3] != 0).sum()), int(pre[4][1])
                    for dim in args.dims:
                        x = torch.randn(nn, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
                        if world == 1:
                            run = lambda: HCSPMM.forward(x, rp, ci, *pre)[0]
                        else:
                            x_loc = x[sg.r0:sg.r1].contiguous()
                            run = lambda: sg.aggregate(x_loc)
                        ms = timeit(run)
                        y = run()
                        if dim not in refs:
                            refs[dim] = y.clone()
                        err = float((y - refs[dim]).norm() / refs[dim].norm())
                        t_tcw, t_dg = tcw, dg
                        if world > 1:
                            t = torch.tensor([ms, err, tcw, dg], device=dev, dtype=torch.float64)
                            tmax, tsum = t.clone(), t.clone()
                            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                            dist.all_reduce(tsum)
                            ms, err, t_tcw, t_dg = float(tmax[0]), float(tmax[1]), int(tsum[2]), int(tsum[3])
                        assert err <= 1e-3, (key, err)
                        rows[dim][key] = {"ms": round(ms, 4), "gflops": round(2.0 * ne * dim / ms / 1e6, 1),
                                          "tc_windows": t_tcw, "dense_groups": t_dg, "rel_err_vs_all_cuda": err}
                        del x, y
                    if world > 1:
                        sg.close()
                        del sg
                    del pre
                if rank == 0:
                    for dim in args.dims:
                        print(json.dumps(rows[dim]), flush=True)
                del refs
                del rp, ci
                torch.cuda.empty_cache()
    HCSPMM.set_dense(False)
    HCSPMM.set_classifier("shipped")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
