"""NVLink peer-read roof for the halo exchange, measured with the library's own pull kernel and with a plain device copy
(torchrun, >= 2 GPUs): contiguous rows (every row of the owner, the best case for the pull), every other row, and the
torch copy kernel reading the peer-mapped buffer.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "hc-spmm_b200")):
    sys.path.insert(0, p)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from hcspmm import peer
    pm = peer.PeerMemory(dev)
    out = {"n_gpus": world}
    for dim in (128, 64, 256):
        rows = (256 << 20) // (dim * 4)
        ptr, ptrs = pm.shared(rows * dim * 4)
        mine = pm.tensor(ptr, (rows, dim))
        mine.normal_()
        dst = torch.empty(rows, dim, device=dev)
        nxt = (rank + 1) % world
        theirs = pm.tensor(ptrs[nxt], (rows, dim))
        table = torch.tensor(ptrs, dtype=torch.int64, device=dev)

        def tm(fn, k=10):
            fn()
            pm.barrier()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(k):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / k], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        rec = {}
        for name, step in (("all_rows", 1), ("every_2nd_row", 2), ("every_4th_row", 4)):
            src_row = torch.arange(0, rows, step, device=dev, dtype=torch.int32)
            m = src_row.numel()
            seg = torch.zeros(world + 1, dtype=torch.int32, device=dev)
            seg[nxt + 1:] = m                                   # every row comes from the next rank
            ms = tm(lambda: peer.halo_pull(table, dim, src_row, seg, world, dst[:m], 0, dim, 1 << nxt, nxt))
            rec[f"pull_{name}_gbs"] = round(m * dim * 4 / ms / 1e6, 1)
        ms = tm(lambda: dst.copy_(theirs))
        rec["torch_copy_kernel_gbs"] = round(rows * dim * 4 / ms / 1e6, 1)
        ms = tm(lambda: dst.copy_(mine))
        rec["local_copy_gbs"] = round(rows * dim * 4 / ms / 1e6, 1)
        out[f"dim{dim}"] = rec
        del mine, theirs
    if rank == 0:
        print(json.dumps(out))
    pm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
