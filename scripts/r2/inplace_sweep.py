"""Segment mode of the peer exchange on one shape (torchrun, >= 2 GPUs): for max_refs T in a list and both operand
storages, the step time (barrier + pull of the rows referenced more than T times + SpMM that reads the rest in place
from the owners' operands), its two phases timed on their own, and parity against the single-GPU aggregation.
One JSON line per configuration on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hc-spmm_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="products")
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--refs", type=int, nargs="+", default=[0, 1, 2, 3, 4, 8])
    ap.add_argument("--operands", nargs="+", default=["fp32", "bf16"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--row-blocks", type=int, nargs="+", default=[1])
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import HCSPMM
    from hcspmm import dist as hd, graphs
    rp, ci, info = graphs.named(args.shape, device=dev)
    n, nnz = info["n"], info["nnz"]
    dim = args.dim or info["dim"]
    x_full = torch.randn(n, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
    HCSPMM.set_classifier("shipped")
    pre_full = HCSPMM.preprocess(ci, rp, n, nnz, (n + 15) // 16)
    y_full = HCSPMM.forward(x_full, rp, ci, *pre_full)[0]
    del pre_full

    def tm(fn, k):
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / k], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t), 4)

    for operand in args.operands:
        for T, rb in [(T_, rb_) for T_ in args.refs for rb_ in args.row_blocks]:
            sg = hd.ShardedGraph(rp, ci, schedule="peer", operand=operand, direct_refs=T, row_blocks=rb)
            x_loc = x_full[sg.r0:sg.r1].contiguous()
            xo = sg.own_rows(dim)
            if xo is not None:
                xo.copy_(x_loc)
                x_loc = xo
            for _ in range(3):
                y = sg.aggregate(x_loc)
            step = tm(lambda: sg.aggregate(x_loc), args.steps)
            ex = tm(lambda: sg.exchange(x_loc), 5)
            op = sg.exchange(x_loc)
            dist.barrier()
            sp = tm(lambda: sg.local_spmm(op), 5)
            y = sg.aggregate(x_loc)
            err = torch.tensor([float((y - y_full[sg.r0:sg.r1]).norm() / y_full[sg.r0:sg.r1].norm())], device=dev, dtype=torch.float64)
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            sg.check()
            d = sg.direct
            if rank == 0:
                print(json.dumps({"shape": args.shape, "dim": dim, "n_gpus": world, "operand": operand, "max_refs": T, "row_blocks": sg.blocks["B"] if sg.blocks else 1,
                                  "halo_fraction_first_block": sg.blocks["first_fraction"] if sg.blocks else 1.0,
                                  "step_ms": step, "exchange_only_ms": ex, "spmm_only_ms": sp, "rel_fro_vs_single_gpu": float(err),
                                  "rows_pulled": (d["pulled_rows"] if d else sg.exchange_rows()),
                                  "rows_in_place": d["rows"] if d else 0, "references_in_place": d["refs"] if d else 0,
                                  "x_in_operand": xo is not None}), flush=True)
            sg.close()
            del sg, x_loc, xo, op, y
            torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
