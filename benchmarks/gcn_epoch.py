#!/usr/bin/env python
"""GCN / GIN epoch time on a BASELINE shape (configs[2] / configs[3]) at 1/2/4/8 GPUs.

  python benchmarks/gcn_epoch.py [--shape products --hidden 128 --feat 100 --classes 47 --layers 2]
  torchrun --nproc-per-node N benchmarks/gcn_epoch.py ...

Row-partitioned training (hcspmm.dist.DistGCN): every rank owns a nnz-balanced range of 16-row windows of A and
the matching rows of X / labels.  A layer is routed like the reference's autograd Functions (GNN_model.py:61-232):
the row-local Update GEMM on the library's TMA + tcgen05 kernel, the Aggregation = halo exchange over NVLink peer
memory + local hybrid SpMM, and the reference's FUSED entry points (forward_fixed32_fused: GCN backward, GIN forward)
where it uses them.  Prints one JSON line: median epoch ms (forward + backward + Adam) over --epochs after --warmup,
max over ranks.  At N > 1 rank 0 also trains the SAME model on the unpartitioned graph (outside the timed epochs) and
the line carries the loss difference, epoch by epoch.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--feat", type=int, default=100)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--schedule", default="auto", choices=["auto", "gather", "slabs", "halo", "peer", "push"])
    ap.add_argument("--slabs", type=int, default=1)
    ap.add_argument("--classifier", default="shipped")
    ap.add_argument("--model", default="gcn", choices=["gcn", "gin"],
                    help="gcn: X' = A (X W) (GNN_model.py:61-162); gin: X' = (A X) W (GNN_model.py:166-232)")
    ap.add_argument("--dense", action="store_true", help="tcgen05 dense super-window plans")
    ap.add_argument("--profile", action="store_true", help="print the CUDA-time breakdown of one epoch (torch.profiler)")
    ap.add_argument("--operand", default="fp32", choices=["fp32", "bf16"], help="N > 1: exchange operand storage")
    ap.add_argument("--direct-refs", type=int, default=0,
                    help="N > 1, peer exchange: remote rows referenced at most this many times are read in place by the SpMM "
                         "(0 = off, default; -1 = auto)")
    ap.add_argument("--no-row-sort", action="store_true", help="no row-sorted CSR copy (A/B of csrc/rowsort.cu)")
    ap.add_argument("--row-blocks", type=int, default=0, help="N > 1, peer exchange: row-block pipeline (0 / 1 = off, default)")
    ap.add_argument("--fp32-matmul", action="store_true",
                    help="Update GEMMs (torch.mm) in full FP32; default TF32 like the reference's stack "
                         "(PyTorch 1.8: allow_tf32 on by default; its fused kernels use wmma TF32, :1809-1837)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import HCSPMM
    from hcspmm import dist as hd, graphs
    torch.backends.cuda.matmul.allow_tf32 = not args.fp32_matmul
    HCSPMM.set_dense(bool(args.dense))
    HCSPMM.set_classifier(args.classifier)
    if args.no_row_sort:
        HCSPMM.set_row_sort(False)
    rp, ci, info = graphs.named(args.shape, device=dev, scale=args.scale)
    g = hd.ShardedGraph(rp, ci, schedule=args.schedule, n_slabs=args.slabs, operand=args.operand,
                        direct_refs=None if args.direct_refs < 0 else args.direct_refs,
                        row_blocks=None if args.row_blocks <= 0 else args.row_blocks)
    # the SAME problem at every N: global features / labels from one seed, then this rank's rows
    # (A is binary and unnormalised like the reference's; features are scaled so the logits start O(1))
    gen = torch.Generator(device=dev).manual_seed(100)
    mean_deg = max(1.0, info["nnz"] / info["n"])
    x_all = torch.randn(info["n"], args.feat, device=dev, generator=gen) / mean_deg ** 2
    y_all = torch.randint(0, args.classes, (info["n"],), device=dev, generator=gen)
    x, y = x_all[g.r0:g.r1].contiguous(), y_all[g.r0:g.r1].contiguous()
    order = {"gcn": "auto", "gin": "aggregate_first"}[args.model]

    # the single-GPU reference of the same training (rank 0, unpartitioned graph), for the loss comparison
    ref_losses = None
    if world > 1:
        if rank == 0:
            g1 = hd.ShardedGraph(rp, ci, single=True)
            m1 = hd.DistGCN(g1, args.feat, args.hidden, args.classes, num_layers=args.layers, seed=0, order=order).to(dev)
            o1 = torch.optim.Adam(m1.parameters(), lr=0.01)
            x_all = m1.pad_features(x_all)
            ref_losses = []
            for ep in range(args.warmup + args.epochs):
                o1.zero_grad()
                l1 = m1.loss(x_all, y_all)
                l1.backward()
                o1.step()
                ref_losses.append(float(l1))
            del g1, m1, o1
        dist.barrier()
    del rp, ci, x_all, y_all
    torch.cuda.empty_cache()

    model = hd.DistGCN(g, args.feat, args.hidden, args.classes, num_layers=args.layers, seed=0, order=order).to(dev)
    x = model.pad_features(x)           # feature layout at the padded width (100 -> 104): once, outside the epochs
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    times, losses, all_losses = [], [], []
    for ep in range(args.warmup + args.epochs):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        opt.zero_grad()
        loss = model.loss(x, y)
        loss.backward()
        model.sync_grads()
        opt.step()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        l = loss.detach().double().reshape(1)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(l)
        all_losses.append(float(l))
        if ep >= args.warmup:
            times.append(float(t))
            losses.append(float(l))
    times.sort()
    g.check()
    if args.profile and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            opt.zero_grad()
            loss = model.loss(x, y)
            loss.backward()
            model.sync_grads()
            opt.step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90), file=sys.stderr)

    # where an epoch goes: exchange and local SpMM of every aggregation width, timed on their own (max over ranks)
    def tm(fn, k=5):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / k], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t), 4)
    phases = {}
    with torch.no_grad():
        for w in sorted({min(wt.shape) for wt in model.weights} if args.model == "gcn" else {wt.shape[0] for wt in model.weights}):
            xx = torch.randn(g.n_local, w, device=dev)
            ex_ms = tm(lambda: g.exchange(xx)) if world > 1 else 0.0
            op = g.exchange(xx)                  # the operand of the latest exchange (segment mode reads the peers' too)
            if world > 1:
                dist.barrier()
            phases[f"width{w}"] = {"exchange_ms": ex_ms,
                                   "spmm_ms": tm(lambda: g.local_spmm(op)),
                                   "aggregate_ms": tm(lambda: g.aggregate(xx))}
    if rank == 0:
        print(json.dumps({"metric": "gcn_epoch_ms", "value": times[len(times) // 2], "unit": "ms", "n_gpus": world,
                          "higher_is_better": False, "scaling": "strong", "min_ms": times[0],
                          "config": {"workload": f"{args.layers}-layer {args.model.upper()}, {args.shape}-shape graph", "dense": bool(args.dense), "nodes": info["n"],
                                     "stored_entries": info["nnz"], "feat": args.feat, "hidden": args.hidden,
                                     "classes": args.classes, "schedule": g.schedule, "slabs": g.n_slabs,
                                     "exchange_rows_vs_allgather": (g.exchange_rows() / max(1, (world - 1) * g.max_rows)) if world > 1 else None,
                                     "classifier": args.classifier, "operand": args.operand,
                                     "row_blocks": None if g.blocks is None else {"blocks": g.blocks["B"], "halo_fraction_first_block": g.blocks["first_fraction"]},
                                     "in_place": None if g.direct is None else {k_: g.direct[k_] for k_ in ("T", "rows", "refs", "pulled_rows", "halo_rows")},
                                     "update_gemm": "HCSPMM.gemm_tf32 (cvt.rna TF32, FP32 accumulate); weight gradients torch.mm " +
                                                    ("fp32" if args.fp32_matmul else "tf32")},
                          "phases": phases, "loss_first": losses[0], "loss_last": losses[-1],
                          "loss_trace": all_losses[:3] + all_losses[-2:],
                          "loss_vs_single_gpu": None if ref_losses is None else {
                              "max_abs_diff": max(abs(a_ - b_) for a_, b_ in zip(all_losses, ref_losses)),
                              "max_rel_diff": max(abs(a_ - b_) / max(abs(b_), 1e-12) for a_, b_ in zip(all_losses, ref_losses)),
                              "epochs_compared": len(ref_losses), "single_gpu_trace": ref_losses[:3] + ref_losses[-2:],
                              "how": "rank 0 trains the same model (same seeds) on the unpartitioned graph; every epoch's loss compared"},
                          "routing": "hcspmm.dist.DistGCN: Update GEMMs on HCSPMM.gemm_tf32 (TMA + tcgen05), aggregations through "
                                     "HCSPMM.forward / forward_fixed32_fused on the exchanged operand"}))
    if world > 1:
        g.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
