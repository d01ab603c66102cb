"""Round 2: measured L2 -> SM gather roof, and hcspmm_loa_reorder timing at growing sizes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi, graphs

dev = torch.device("cuda", 0)
res = {"l2": [], "loa": []}
for rf in (256, 512):
    for mb in (16, 32, 64, 96, 256, 1024):
        bw = capi.l2_gather_bandwidth(dev, rf, mb)
        res["l2"].append({"row_floats": rf, "resident_mb": mb, "gbs": bw})
        print(f"l2 gather row_floats {rf} buffer {mb} MB: {bw:.0f} GB/s")
if "--loa" in sys.argv:
    for n, e in ((20000, 500000), (80000, 2000000), (320000, 8000000)):
        rp, ci = graphs.rmat(n, e, seed=5)
        rp, ci = rp.to(dev), ci.to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        perm, sizes, nf = capi.loa_reorder(rp, ci)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res["loa"].append({"n": n, "nnz": int(ci.numel()), "seconds": dt, "blocks": int(sizes.numel()), "full": nf})
        print(f"loa n {n} nnz {ci.numel()}: {dt:.2f} s, {sizes.numel()} blocks ({nf} full)")
        if dt > 40:
            break
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_l2_loa.json"), "w"), indent=1)
