"""Staged gather (knob "staged") against the register-ring balanced kernel: products shape at the widths the kernel
serves, a 1/8 row shard, and a smaller R-MAT."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
import HCSPMM
from hcspmm import graphs, partition
dev = torch.device("cuda", 0)


def t(fn, n=20):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


rp, ci, info = graphs.named("products", device=dev)
n = info["n"]
pre = HCSPMM.preprocess(ci, rp, n, ci.numel(), (n + 15) // 16)
rec = {"shape": "products"}
for dim in (128, 104, 96):
    x = torch.randn(n, dim, device=dev)
    for staged in (0, 1, 0, 1):
        HCSPMM.set_tuning("staged", staged)
        rec.setdefault(f"dim{dim}_staged{staged}", []).append(t(lambda: HCSPMM.forward(x, rp, ci, *pre)))
    HCSPMM.set_tuning("staged", 1); y1 = HCSPMM.forward(x, rp, ci, *pre)[0]
    HCSPMM.set_tuning("staged", 0); y0 = HCSPMM.forward(x, rp, ci, *pre)[0]
    rec[f"dim{dim}_rel_diff"] = float((y1 - y0).norm() / y0.norm())
print(json.dumps(rec), flush=True)
# a 1/8 row shard (what a rank of 8 runs), rectangular
cuts = partition.window_cuts(rp, 8)
rp_l, ci_l = partition.local_shard(rp, ci, cuts[3], cuts[4])
nl = rp_l.numel() - 1
pre_l = HCSPMM.preprocess(ci_l, rp_l, nl, ci_l.numel(), (nl + 15) // 16)
x = torch.randn(n, 128, device=dev); out = torch.empty(nl, 128, device=dev)
rec = {"shape": "products 1/8 shard"}
for staged in (0, 1, 0, 1):
    HCSPMM.set_tuning("staged", staged)
    rec.setdefault(f"staged{staged}", []).append(t(lambda: HCSPMM.spmm_strided(x, rp_l, ci_l, *pre_l[:4], out, False, *pre_l[4:6])))
print(json.dumps(rec), flush=True)
HCSPMM.set_tuning("staged", 0)
