"""Round 2: timing of the dense super-window kernels (cp.async vs TMA gather4) and of fused vs unfused
Aggregation + Update on the proteins shape (GIN, hidden 256) and on banded graphs."""
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


out = []
cases = [("proteins", None)]
for shape, _ in cases:
    rp, ci, info = graphs.named(shape, device=dev)
    n, nnz = info["n"], info["nnz"]
    bp, etc, etr, ht = capi.preprocess(ci, rp, "b200")
    plan = capi.DensePlan(rp, ci, etr, ht, min_reuse=2.0)
    aux = capi.GraphAux(rp, ci, ht, plan)
    print(f"{shape}: n {n} nnz {nnz} dense super-windows {plan.n_dense} condensed columns {plan.total_cols} plan_full {aux.plan_full}")
    for dim, hidden in ((256, 256), (128, 128), (64, 64)):
        x = torch.randn(n, dim, device=dev)
        w = torch.randn(dim, hidden, device=dev)
        rec = {"shape": shape, "dim": dim, "hidden": hidden, "n_dense": plan.n_dense, "total_cols": plan.total_cols}
        for name, tma, ws in (("dense.cu single-role", 0, 0), ("dense.cu warp-specialised", 0, 1),
                              ("dense_tma.cu cp.async gather", 2, 1), ("dense_tma.cu tma gather4", 3, 1)):
            capi.set_tuning("dense_tma", tma); capi.set_tuning("dense_ws", ws)
            rec["spmm_ms " + name] = t(lambda: capi.spmm_aux(x, rp, ci, bp, etc, etr, ht, aux))
        for tma in (0, 1, 2, 3):
            capi.set_tuning("dense_tma", tma)
            for name, fuse in (("unfused", 0), ("fused", 1)):
                capi.set_tuning("fuse_update", fuse)
                rec[f"spmm_gemm_ms dense_tma={tma} {name}"] = t(lambda: capi.spmm_gemm_aux(x, rp, ci, bp, etc, etr, ht, w, aux))
        capi.set_tuning("fuse_update", 1); capi.set_tuning("dense_tma", 1)
        rec["gemm_only_ms"] = t(lambda: capi.gemm_tf32(x, w))
        aux0 = capi.GraphAux(rp, ci, ht)
        rec["spmm_ms cuda cores (no plan)"] = t(lambda: capi.spmm_aux(x, rp, ci, bp, etc, etr, ht, aux0))
        flops_exec = 2.0 * 128 * plan.total_cols * dim
        best = min(v for k, v in rec.items() if k.startswith("spmm_ms dense"))
        rec["executed_tflops best dense"] = flops_exec / best / 1e9
        rec["useful_gflops best dense"] = 2.0 * nnz * dim / best / 1e6
        print(json.dumps(rec))
        out.append(rec)
print("umma err", capi.lib().hcspmm_debug_umma_error())
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_dense_probe.json"), "w"), indent=1)
