"""What part of the degree-order gain on low-degree graphs is ROW order (which rows share an item of the balanced kernel:
invisible to the caller, the kernel could keep a row-sorted copy of the CSR) and what part is COLUMN order (which rows
of X are neighbours in memory: needs relabelled vertex ids)?  Shapes: products (dim 128) and a 1/8 shard-sized R-MAT."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi, graphs

dev = torch.device("cuda", 0)


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def measure(rp, ci, dim, x_rows=None):
    n = rp.numel() - 1
    bp, etc, etr, ht = capi.preprocess(ci, rp, "shipped")
    aux = capi.GraphAux(rp, ci, ht)
    x = torch.randn(x_rows or n, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    return round(t(lambda: capi.spmm_aux(x, rp, ci, bp, etc, etr, ht, aux)), 4)


def permute_rows(rp, ci, perm):
    """new row i = old row perm[i]; column ids unchanged."""
    deg = (rp[1:] - rp[:-1]).long()
    nd = deg[perm]
    nrp = torch.zeros_like(rp)
    nrp[1:] = torch.cumsum(nd, 0).to(rp.dtype)
    start = rp[:-1].long()[perm]
    idx = torch.repeat_interleave(start - nrp[:-1].long(), nd) + torch.arange(int(nd.sum()), device=rp.device)
    return nrp, ci[idx].contiguous()


def permute_cols(rp, ci, perm):
    """column c -> inv[c] (X would be stored in that order); rows unchanged, entries re-sorted inside each row."""
    n = rp.numel() - 1
    inv = torch.empty(n, dtype=torch.int64, device=rp.device)
    inv[perm] = torch.arange(n, device=rp.device)
    rows = torch.repeat_interleave(torch.arange(n, device=rp.device), (rp[1:] - rp[:-1]).long())
    key = rows * n + inv[ci.long()]
    key, _ = torch.sort(key)
    return rp, (key % n).to(torch.int32).contiguous()


for name, dim in (("products", 128),):
    rp, ci, info = graphs.named(name, device=dev)
    deg = (rp[1:] - rp[:-1]).long()
    by_deg = torch.argsort(-deg, stable=True)
    rec = {"shape": name, "dim": dim, "as_generated": measure(rp, ci, dim)}
    r2, c2 = permute_rows(rp, ci, by_deg)
    rec["rows_by_degree"] = measure(r2, c2, dim)
    r3, c3 = permute_cols(rp, ci, by_deg)
    rec["columns_by_degree"] = measure(r3, c3, dim)
    r4, c4 = permute_cols(r2, c2, by_deg)
    rec["rows_and_columns_by_degree"] = measure(r4, c4, dim)
    # rows bucketed by degree class only (log2 buckets, original order inside): what a cheap counting sort would give
    cls = torch.floor(torch.log2(deg.clamp(min=1).float())).long()
    by_cls = torch.argsort(-cls, stable=True)
    r5, c5 = permute_rows(rp, ci, by_cls)
    rec["rows_by_log2_degree_class"] = measure(r5, c5, dim)
    print(json.dumps(rec), flush=True)
