#!/bin/bash
# round 2, 1 GPU: row-sorted CSR copy -- tests, products / sweep timing, full GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rowsort.py tests/test_gpu_peer.py -x -q -m gpu > gpurun_out/r2_c17_tests_new.log 2>&1; echo "new tests rc=$?"; tail -8 gpurun_out/r2_c17_tests_new.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c17_tests_all.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/r2_c17_tests_all.log
python - <<'PY'
import sys, json, torch
sys.path[:0] = [".", "hc-spmm_b200"]
import HCSPMM
from hcspmm import graphs
dev = torch.device("cuda", 0)
def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)
for shape, dims in (("products", (128, 48, 104, 256)), ("reddit", (256,)), ("proteins", (256,))):
    rp, ci, info = graphs.named(shape, device=dev)
    n = info["n"]
    for sort in (False, True):
        if shape != "products" and sort:
            continue
        HCSPMM.set_row_sort(sort)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pre = HCSPMM.preprocess(ci, rp, n, ci.numel(), (n + 15) // 16)
        torch.cuda.synchronize(); a.record()
        pre = HCSPMM.preprocess(ci, rp, n, ci.numel(), (n + 15) // 16)
        b.record(); torch.cuda.synchronize()
        rec = {"shape": shape, "row_sort": sort, "sorted_copy": int(pre[4][15]) > 0, "preprocess_ms": round(a.elapsed_time(b), 3)}
        for dim in dims:
            x = torch.randn(n, dim, device=dev)
            rec[f"dim{dim}_ms"] = t(lambda: HCSPMM.forward(x, rp, ci, *pre))
            if dim % 8 == 0:
                xb = x.to(torch.bfloat16); out = torch.empty(n, dim, device=dev)
                rec[f"dim{dim}_bf16_stored_ms"] = t(lambda: HCSPMM.spmm_bf16(xb, rp, ci, *pre[:4], out, False, *pre[4:6]))
        print(json.dumps(rec), flush=True)
    HCSPMM.set_row_sort(True)
PY
