"""Round 2: L2 residency hints of the balanced kernel (tagged column ids, evict_last for the hottest rows that fit the
budget, evict_first for the rest) against round 1's evict_last-for-every-row, per shape / width / budget."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi, graphs

dev = torch.device("cuda", 0)
capi.set_tuning("l2_hot_min_row", 0)      # measure the hints at every width


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = []
for shape, dims in (("products", (128, 64, 256)), ("reddit", (256, 512, 128))):
    rp, ci, info = graphs.named(shape, device=dev)
    n = info["n"]
    pre = capi.preprocess(ci, rp, "shipped")
    plain = capi.GraphAux(rp, ci, pre[3])
    tagged = capi.GraphAux(rp, ci, pre[3], tag_columns=True, n_cols=n)
    cls = (tagged.tagged.view(torch.int32) >> 29) & 7
    hist = torch.bincount(cls.long(), minlength=8).tolist()
    print(shape, "entries per hotness class 0..7:", hist, flush=True)
    for dim in dims:
        x = torch.randn(n, dim, device=dev)
        rec = {"shape": shape, "dim": dim, "x_mb": n * dim * 4 / 1e6}
        rec["evict_last for all (round 1)"] = t(lambda: capi.spmm_aux(x, rp, ci, *pre, plain))
        y0 = capi.spmm_aux(x, rp, ci, *pre, plain)
        for mb in (24, 48, 72, 96, 120):
            capi.set_tuning("l2_hot_mb", mb)
            rec[f"hints, budget {mb} MB"] = t(lambda: capi.spmm_aux(x, rp, ci, *pre, tagged))
            assert torch.equal(capi.spmm_aux(x, rp, ci, *pre, tagged), y0), "hints changed the result"
        capi.set_tuning("l2_hot_mb", 72)
        old = capi.set_tuning("occupancy3", 0)
        rec["hints 72 MB, 2 CTAs/SM build"] = t(lambda: capi.spmm_aux(x, rp, ci, *pre, tagged))
        rec["evict_last for all, 2 CTAs/SM build"] = t(lambda: capi.spmm_aux(x, rp, ci, *pre, plain))
        capi.set_tuning("occupancy3", old)
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)
        out.append(rec)
    del rp, ci, pre, plain, tagged
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_hint_probe.json"), "w"), indent=1)
