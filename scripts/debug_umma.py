import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch, numpy as np
from hcspmm import capi
torch.manual_seed(0)
capi.set_tuning("umma_gemm", 1)
def run(m, k, n, a=None, b=None, tag=""):
    a = torch.randn(m, k, device="cuda") if a is None else a
    b = torch.randn(k, n, device="cuda") if b is None else b
    out = capi.gemm_tf32(a, b)
    torch.cuda.synchronize()
    err = capi.lib().hcspmm_debug_umma_error()
    ref = a.double() @ b.double()
    rel = float((out.double() - ref).norm() / ref.norm())
    print(f"[{tag}] m{m} k{k} n{n}: err_flag={err} rel={rel:.3e} |out|={float(out.norm()):.3e} |ref|={float(ref.norm()):.3e} nonzero={int((out != 0).sum())}/{out.numel()}")
    return out, ref
# identity A: out should equal B's first rows
m, k, n = 128, 32, 32
a = torch.zeros(m, k, device="cuda"); a[:32, :32] = torch.eye(32, device="cuda")
b = torch.arange(k * n, device="cuda", dtype=torch.float32).reshape(k, n)
out, ref = run(m, k, n, a, b, "eyeA")
print("out[0:4,0:8]\n", out[0:4, 0:8].cpu().numpy()); print("ref[0:4,0:8]\n", ref[0:4, 0:8].cpu().numpy())
print("out[32:34,0:8]\n", out[32:34, 0:8].cpu().numpy())
# ones
out, ref = run(128, 32, 32, torch.ones(128, 32, device="cuda"), torch.ones(32, 32, device="cuda"), "ones")
print("out[0,0:8]", out[0, 0:8].cpu().numpy(), "out[127,0:8]", out[127, 0:8].cpu().numpy())
# A = ones, B = row index  -> each out = sum_k k * 1... col pattern
b = torch.arange(32, device="cuda", dtype=torch.float32).reshape(32, 1).repeat(1, 32)
out, ref = run(128, 32, 32, torch.ones(128, 32, device="cuda"), b, "onesA_Brow")
print("out[0,0:8]", out[0, 0:8].cpu().numpy(), "ref", ref[0, 0:8].cpu().numpy())
b = torch.arange(32, device="cuda", dtype=torch.float32).reshape(1, 32).repeat(32, 1)
out, ref = run(128, 32, 32, torch.ones(128, 32, device="cuda"), b, "onesA_Bcol")
print("out[0,0:32]", out[0, 0:32].cpu().numpy()); print("ref", ref[0, 0:32].cpu().numpy())
a = torch.arange(128, device="cuda", dtype=torch.float32).reshape(128, 1).repeat(1, 32)
out, ref = run(128, 32, 32, a, torch.ones(32, 32, device="cuda"), "Arow_onesB")
print("out[:8,0]", out[:8, 0].cpu().numpy(), "ref", ref[:8, 0].cpu().numpy())
a = torch.zeros(128, 32, device="cuda"); a[:, 5] = 1.0
b = torch.zeros(32, 32, device="cuda"); b[5, :] = torch.arange(32, device="cuda", dtype=torch.float32)
out, ref = run(128, 32, 32, a, b, "k5")
print("out[0,0:32]", out[0, 0:32].cpu().numpy())
for shp in [(128, 64, 256), (1000, 128, 128), (513, 100, 48), (300, 36, 16)]:
    run(*shp, tag="rand")
