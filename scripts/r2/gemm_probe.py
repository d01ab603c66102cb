"""Round 2: correctness + timing of the TMA/tcgen05 Update GEMM against the other kernels and cuBLAS."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import numpy as np
import torch
import oracle
from hcspmm import capi


def t(fn, n=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


print("== correctness (vs oracle cvt.rna TF32 GEMM)")
for (m, k, n) in [(128, 32, 32), (1000, 128, 128), (513, 100, 48), (4096, 256, 256), (300, 36, 16), (2000, 128, 320), (77, 8, 4), (5000, 128, 128)]:
    g = torch.Generator().manual_seed(m + k + n)
    a, b = torch.randn(m, k, generator=g), torch.randn(k, n, generator=g)
    want = oracle.gemm(a.numpy(), b.numpy(), tf32=True).astype(np.float64)
    row = {}
    for name, um, rnd in (("mma.sync", 0, 1), ("tcgen05-reg", 1, 1), ("tma-round", 2, 1), ("tma-tf32map", 2, 0)):
        capi.set_tuning("umma_gemm", um); capi.set_tuning("gemm_round", rnd)
        got = capi.gemm_tf32(a.cuda(), b.cuda()).cpu().numpy()
        row[name] = rel(got, want)
    print(m, k, n, {k_: f"{v:.2e}" for k_, v in row.items()}, "err flag", capi.lib().hcspmm_debug_umma_error())

print("== timing")
out = []
for (m, k, n) in [(2449029, 128, 128), (2449029, 100, 128), (2449029, 128, 48), (232965, 256, 256), (132534, 256, 256), (1048576, 32, 32), (2449029, 64, 64)]:
    a = torch.randn(m, k, device="cuda"); b = torch.randn(k, n, device="cuda")
    res = {}
    for name, um, rnd, st in (("mma.sync", 0, 1, 0), ("tcgen05-reg", 1, 1, 0), ("tma-round", 2, 1, 0), ("tma-tf32map", 2, 0, 0),
                              ("tma-round-s2", 2, 1, 2), ("tma-round-s3", 2, 1, 3)):
        capi.set_tuning("umma_gemm", um); capi.set_tuning("gemm_round", rnd); capi.set_tuning("gemm_stages", st)
        res[name] = t(lambda: capi.gemm_tf32(a, b))
    capi.set_tuning("gemm_stages", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    res["cublas-tf32"] = t(lambda: torch.mm(a, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    gb = (m * k + m * n) * 4 / 1e9
    res = {k_: round(v, 4) for k_, v in res.items()}
    best = res["tma-round"]
    print(f"m{m} k{k} n{n}: {res} | tma-round {gb / best * 1e3:.0f} GB/s = {gb / best * 1e3 / 6446.3:.2f} of HBM copy peak")
    out.append({"m": m, "k": k, "n": n, "ms": res, "bytes": gb * 1e9, "hbm_frac_tma_round": gb / best * 1e3 / 6446.3})
print("umma err flag", capi.lib().hcspmm_debug_umma_error())
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_gemm_timing.json"), "w"), indent=1)
