import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import capi
def t(fn, n=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (m, k, n) in [(2449029, 128, 128), (2449029, 100, 128), (232965, 256, 256), (2449029, 128, 48), (132534, 256, 256), (1048576, 32, 32)]:
    a = torch.randn(m, k, device="cuda"); b = torch.randn(k, n, device="cuda")
    res = {}
    for umma in (0, 1):
        capi.set_tuning("umma_gemm", umma)
        res[umma] = t(lambda: capi.gemm_tf32(a, b))
    torch.backends.cuda.matmul.allow_tf32 = True
    tt = t(lambda: torch.mm(a, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    tf = t(lambda: torch.mm(a, b))
    gb = (m * k + m * n) * 4 / 1e9
    print(f"m{m} k{k} n{n}: mma.sync {res[0]:.3f} ms | tcgen05 {res[1]:.3f} ms ({2*m*k*n/res[1]/1e9:.0f} TFLOP/s... {gb/res[1]*1e3:.0f} GB/s) | torch tf32 {tt:.3f} | torch fp32 {tf:.3f}")
print("umma err flag", capi.lib().hcspmm_debug_umma_error())
