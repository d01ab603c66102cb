"""Round 2 profiling targets: each scenario launches its kernel 4 times (ncu: -k regex:<kernel> -s 2 -c 1).
  reddit_spmm | products_spmm | gemm_products | dense_proteins | fused_proteins | bench_step"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi, graphs

what = sys.argv[1]
dev = torch.device("cuda", 0)
if what in ("reddit_spmm", "products_spmm", "products_spmm_degree", "products_spmm_staged"):
    shape = what.split("_")[0]
    if what.endswith("staged"):          # shared-memory staged gather (knob "staged")
        capi.set_tuning("staged", 1)
    rp, ci, info = graphs.named(shape, device=dev)
    if what.endswith("degree"):          # descending-degree relabelling (scripts/r2/locality.py)
        deg = (rp[1:] - rp[:-1]).long()
        rp, ci = capi.relabel(rp, ci, torch.argsort(-deg, stable=True).to(torch.int32))
    dim = info["dim"]
    pre = capi.preprocess(ci, rp, "shipped")
    aux = capi.GraphAux(rp, ci, pre[3])
    x = torch.randn(info["n"], dim, device=dev)
    for _ in range(4):
        capi.spmm_aux(x, rp, ci, *pre, aux)
elif what == "gemm_products":
    a, b = torch.randn(2449029, 128, device=dev), torch.randn(128, 128, device=dev)
    for _ in range(4):
        capi.gemm_tf32(a, b)
elif what in ("dense_proteins", "fused_proteins"):
    rp, ci, info = graphs.named("proteins", device=dev)
    pre = capi.preprocess(ci, rp, "b200")
    plan = capi.DensePlan(rp, ci, pre[2], pre[3], min_reuse=2.0)
    aux = capi.GraphAux(rp, ci, pre[3], plan)
    x, w = torch.randn(info["n"], 256, device=dev), torch.randn(256, 256, device=dev)
    for _ in range(4):
        if what == "dense_proteins":
            capi.spmm_aux(x, rp, ci, *pre, aux)
        else:
            capi.spmm_gemm_aux(x, rp, ci, *pre, w, aux)
torch.cuda.synchronize()
print("done", what)
