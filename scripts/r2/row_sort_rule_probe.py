"""Where does the row-sorted copy pay?  Register-ring balanced kernel with / without it on graphs outside the rule
(8 <= mean row < 64): Reddit shape (mean 492), proteins shape (uniform 296), R-MATs of mean 16 / 64 / 128."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi, graphs
dev = torch.device("cuda", 0)


def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


cases = [("reddit", None, 256), ("proteins", None, 256), ("rmat16", (1_000_000, 16_000_000), 128),
         ("rmat64", (500_000, 32_000_000), 128), ("rmat128", (400_000, 51_200_000), 128), ("rmat128", (400_000, 51_200_000), 256)]
for name, spec, dim in cases:
    if spec is None:
        rp, ci, info = graphs.named(name, device=dev)
    else:
        rp, ci = graphs.rmat(spec[0], spec[1], seed=3, device=dev)
    n = rp.numel() - 1
    pre = capi.preprocess(ci, rp, "shipped")
    x = torch.randn(n, dim, device=dev)
    rec = {"graph": name, "n": n, "nnz": int(ci.numel()), "mean_row": round(ci.numel() / n, 1), "dim": dim}
    for sort in (False, True, False, True):
        aux = capi.GraphAux(rp, ci, pre[3], row_sort=sort)
        rec.setdefault("sorted" if sort else "in_place", []).append(t(lambda: capi.spmm_aux(x, rp, ci, *pre, aux)))
        del aux
    print(json.dumps(rec), flush=True)
    del rp, ci, pre, x
    torch.cuda.empty_cache()
