"""A/B of two builds of libhcspmm on the same box: hcspmm_spmm (no per-graph products) on a named shape."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200")]
import torch
from hcspmm import graphs
dev = torch.device("cuda", 0)
libs = {"old": os.path.join(ROOT, "scripts/debug/_ab/libhcspmm_old.so"), "new": os.path.join(ROOT, "hc-spmm_b200/lib/libhcspmm.so")}
vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
for shape, dims in (("products", (128, 48)), ("reddit", (256,))):
    rp, ci, info = graphs.named(shape, device=dev)
    n, nnz = info["n"], info["nnz"]
    rec = {"shape": shape}
    for dim in dims:
        x = torch.randn(n, dim, device=dev); y = torch.empty(n, dim, device=dev)
        for rep in range(2):
            for name, path in libs.items():
                L = ctypes.CDLL(path)
                L.hcspmm_spmm.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp, i32, i64, i32, ctypes.c_int, ctypes.c_int, vp, i64, vp]
                def run():
                    rc = L.hcspmm_spmm(x.data_ptr(), dim, n, rp.data_ptr(), ci.data_ptr(), None, None, None, None, n, nnz, dim, 2, 0, y.data_ptr(), dim, None)
                    assert rc == 0, rc
                for _ in range(3): run()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(10): run()
                b.record(); torch.cuda.synchronize()
                rec[f"dim{dim}_{name}_{rep}"] = round(a.elapsed_time(b) / 10, 4)
    print(json.dumps(rec), flush=True)
