"""GPU tests of the drop-in `HCSPMM` torch extension: the 18 names of the reference's pybind module
(hybrid_all.cpp:500-525) with the reference's positional signatures, called the way GNN_model.py
calls them, checked against the CPU oracle; plus the recompiled UNMODIFIED reference extension
(oracle/_ref/HCSPMM_ref.so) run beside ours inside its validity envelope."""
import importlib.machinery
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import rel_fro, small_graphs

pytestmark = pytest.mark.gpu
GRAPHS = small_graphs()
REF_SO = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "HCSPMM_ref.so")

REFERENCE_NAMES = [
    "preprocess", "forward", "forward_more", "forward_fixed32", "forward_fixed32_fused", "forward_final_fused",
    "forward_fixed64", "forward_fixed64_fused", "forward_final_fused_64", "forward_GIN_final_fused",
    "backward", "backward_fixed32", "backward_fixed32_fused", "backward_final_fused", "backward_fixed64",
    "backward_fixed64_fused", "backward_final_fused_64", "backward_GIN_final_fused"]


@pytest.fixture(scope="module")
def H():
    import HCSPMM
    return HCSPMM


@pytest.fixture(scope="module")
def REF():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/HCSPMM_ref.so not built")
    loader = importlib.machinery.ExtensionFileLoader("HCSPMM_ref", REF_SO)
    spec = importlib.util.spec_from_loader("HCSPMM_ref", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def prep(H, rp, ci, classifier="shipped"):
    old = H.set_classifier(classifier)
    try:
        n = rp.size - 1
        return H.preprocess(dev(ci), dev(rp), n, ci.size, (n + 15) // 16)
    finally:
        H.set_classifier(old)


def test_module_exports_reference_names(H):
    for name in REFERENCE_NAMES:
        assert callable(getattr(H, name)), name


@pytest.mark.parametrize("name", ["ring3_256", "rmat_1000", "holes_777"])
def test_preprocess_and_forward_like_gnn_model(H, name):
    rp, ci = GRAPHS[name]
    n = rp.size - 1
    pre = prep(H, rp, ci)
    assert len(pre) == 6 and all(t.dtype == torch.int32 for t in pre)
    # the four reference arrays and the per-graph blob live on the device; row_nzr is the HOST header of the blob
    assert all(t.is_cuda for t in pre[:4]) and pre[5].is_cuda and not pre[4].is_cuda
    want = oracle.preprocess(ci, rp, oracle.MODE_SHIPPED)
    for got, w in zip(pre[:4], want):
        assert np.array_equal(got.cpu().numpy(), w)
    x = torch.randn(n, 32, generator=torch.Generator().manual_seed(0))
    for fn in (H.forward, H.forward_fixed32, H.forward_more, H.forward_fixed64, H.backward, H.backward_fixed32):
        out = fn(x.cuda(), dev(rp), dev(ci), *pre)
        assert isinstance(out, list) and len(out) == 1 and out[0].shape == (n, 32)
        assert rel_fro(out[0].cpu().numpy(), oracle.spmm(rp, ci, x.numpy(), precision=1)) <= 1e-5


def test_fused_entry_points(H):
    rp, ci = GRAPHS["rmat_1000"]
    n = 1000
    pre = prep(H, rp, ci)
    g = torch.Generator().manual_seed(1)
    for dim, hidden in ((32, 32), (128, 128), (22, 32), (100, 47)):
        x, w = torch.randn(n, dim, generator=g), torch.randn(dim, hidden, generator=g)
        z_want = oracle.spmm(rp, ci, x.numpy(), precision=1)
        o_want = oracle.gemm(z_want, w.numpy(), tf32=True)
        for fn in (H.forward_fixed32_fused, H.forward_fixed64_fused, H.forward_GIN_final_fused,
                   H.backward_fixed32_fused):
            out, z = fn(x.cuda(), dev(rp), dev(ci), *pre, w.cuda())
            assert out.shape == (n, hidden) and z.shape == (n, dim)
            assert rel_fro(z.cpu().numpy(), z_want) <= 1e-5
            assert rel_fro(out.cpu().numpy(), o_want) <= 1e-4
        # final_fused writes the caller's buffer in place and returns it (hybrid_all_kernel.cu:748,756)
        buf = torch.full((n, hidden), 7.0, device="cuda")
        out, z = H.forward_final_fused(x.cuda(), dev(rp), dev(ci), *pre, w.cuda(), buf)
        assert out.data_ptr() == buf.data_ptr()
        assert rel_fro(buf.cpu().numpy(), o_want) <= 1e-4


def test_weights_view_is_honoured_unless_bug_compat(H):
    """GNN_model.py:98,120 pass weights.transpose(0,1); the reference reads the view's raw memory
    (SURVEY.md 3.4-1).  Default: the view's values; bug_compat: the reference's reading."""
    rp, ci = GRAPHS["ring3_256"]
    pre = prep(H, rp, ci)
    g = torch.Generator().manual_seed(2)
    x, w = torch.randn(256, 32, generator=g), torch.randn(32, 32, generator=g)
    z = oracle.spmm(rp, ci, x.numpy(), precision=1)
    out = H.forward_fixed32_fused(x.cuda(), dev(rp), dev(ci), *pre, w.cuda().transpose(0, 1))[0]
    assert rel_fro(out.cpu().numpy(), oracle.gemm(z, w.numpy().T.copy(), tf32=True)) <= 1e-4
    H.set_bug_compat(True)
    try:
        out = H.forward_fixed32_fused(x.cuda(), dev(rp), dev(ci), *pre, w.cuda().transpose(0, 1))[0]
    finally:
        H.set_bug_compat(False)
    assert rel_fro(out.cpu().numpy(), oracle.gemm(z, w.numpy(), tf32=True)) <= 1e-4


def test_error_behaviour(H):
    rp, ci = GRAPHS["ring3_256"]
    pre = prep(H, rp, ci)
    x = torch.randn(256, 32)
    with pytest.raises(RuntimeError, match="CUDA tensor"):       # same convention as CHECK_CUDA
        H.forward(x, dev(rp), dev(ci), *pre)
    with pytest.raises(RuntimeError, match="contiguous"):        # CHECK_CONTIGUOUS
        H.forward(torch.randn(32, 256).cuda().t(), dev(rp), dev(ci), *pre)
    with pytest.raises(RuntimeError, match="int32"):
        H.forward(x.cuda(), dev(rp).long(), dev(ci), *pre)
    with pytest.raises(RuntimeError, match="float32"):
        H.forward(x.cuda().double(), dev(rp), dev(ci), *pre)
    with pytest.raises(RuntimeError, match="num_row_windows"):
        H.preprocess(dev(ci), dev(rp), 256, ci.size, 3)


def test_non_default_stream_and_accumulate(H):
    rp, ci = GRAPHS["rmat_1000"]
    pre = prep(H, rp, ci)
    x = torch.randn(1000, 64, generator=torch.Generator().manual_seed(3))
    want = oracle.spmm(rp, ci, x.numpy(), precision=1)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        xd = x.cuda()
        out = H.forward(xd, dev(rp), dev(ci), *pre)[0]
        acc = torch.ones(1000, 64, device="cuda")
        H.spmm_accumulate(xd, dev(rp), dev(ci), *pre[:4], acc, True)
    s.synchronize()
    assert rel_fro(out.cpu().numpy(), want) <= 1e-5
    assert rel_fro(acc.cpu().numpy(), want + 1.0) <= 1e-5


# ---- the unmodified reference, recompiled for sm_100, inside its envelope --------------------
def test_reference_preprocess_matches(H, REF):
    """<= 62 edges and <= 24 distinct columns per window, N % 16 == 0 (hybrid_all_kernel.cu:23,26)."""
    for name in ("ring3_256", "band2_320"):
        rp, ci = GRAPHS[name]
        n = rp.size - 1
        ref = REF.preprocess(dev(ci), dev(rp), n, ci.size, n // 16)
        torch.cuda.synchronize()
        ours = prep(H, rp, ci)
        orc = oracle.preprocess(ci, rp, oracle.MODE_SHIPPED)
        for r, o, c, what in zip(ref[:4], ours[:4], orc, ("blockPartition", "edgeToColumn", "edgeToRow", "hybrid_type")):
            assert np.array_equal(r.cpu().numpy(), c), f"oracle != reference: {what} on {name}"
            assert np.array_equal(r.cpu().numpy(), o.cpu().numpy()), f"ours != reference: {what} on {name}"


def test_reference_spmm_cuda_path_matches(H, REF):
    rp, ci = GRAPHS["ring3_256"]
    n = 256
    ref_pre = REF.preprocess(dev(ci), dev(rp), n, ci.size, n // 16)
    pre = prep(H, rp, ci)
    g = torch.Generator().manual_seed(4)
    for dim, ref_fn, our_fn in ((32, REF.forward_fixed32, H.forward_fixed32), (32, REF.forward, H.forward),
                                (16, REF.forward, H.forward), (48, REF.forward, H.forward),
                                (96, REF.forward, H.forward)):
        x = torch.randn(n, dim, generator=g).cuda()
        r = ref_fn(x, dev(rp), dev(ci), *ref_pre)[0]
        o = our_fn(x, dev(rp), dev(ci), *pre)[0]
        torch.cuda.synchronize()
        assert rel_fro(o.cpu().numpy(), r.cpu().numpy()) <= 1e-6
        assert rel_fro(r.cpu().numpy(), oracle.spmm(rp, ci, x.cpu().numpy(), precision=1)) <= 1e-6


def test_reference_spmm_tensor_core_path_matches(H, REF):
    """Hand-forced hybrid_type = 1 on the banded fixture (<= 20 distinct columns per window)."""
    rp, ci = GRAPHS["band2_320"]
    n = 320
    ref_pre = list(REF.preprocess(dev(ci), dev(rp), n, ci.size, n // 16))
    ref_pre[3] = torch.ones(n // 16, dtype=torch.int32, device="cuda")
    pre = list(prep(H, rp, ci))
    pre[3] = torch.ones(n // 16, dtype=torch.int32, device="cuda")
    g = torch.Generator().manual_seed(5)
    for dim in (16, 32, 48):
        x = torch.randn(n, dim, generator=g).cuda()
        r = REF.forward(x, dev(rp), dev(ci), *ref_pre)[0]
        o = H.forward(x, dev(rp), dev(ci), *pre)[0]
        torch.cuda.synchronize()
        fp32 = oracle.spmm(rp, ci, x.cpu().numpy(), precision=1)
        assert rel_fro(r.cpu().numpy(), fp32) <= 1e-3          # the reference itself, TF32
        assert rel_fro(o.cpu().numpy(), r.cpu().numpy()) <= 2e-5
        assert rel_fro(r.cpu().numpy(), oracle.spmm(rp, ci, x.cpu().numpy(), hybrid_type=np.ones(20, np.int32))) <= 2e-5


def test_reference_fused_matches(H, REF):
    rp, ci = GRAPHS["ring3_256"]
    n = 256
    ref_pre = REF.preprocess(dev(ci), dev(rp), n, ci.size, n // 16)
    pre = prep(H, rp, ci)
    g = torch.Generator().manual_seed(6)
    x, w = torch.randn(n, 32, generator=g).cuda(), torch.randn(32, 32, generator=g).cuda()
    r_out, r_z = REF.forward_fixed32_fused(x, dev(rp), dev(ci), *ref_pre, w)
    o_out, o_z = H.forward_fixed32_fused(x, dev(rp), dev(ci), *pre, w)
    torch.cuda.synchronize()
    assert rel_fro(o_z.cpu().numpy(), r_z.cpu().numpy()) <= 1e-6
    assert rel_fro(o_out.cpu().numpy(), r_out.cpu().numpy()) <= 1e-4


def test_dense_plan_travels_in_row_nzr_col_nzr(H):
    """set_dense(True): preprocess() returns the tcgen05 dense plan in the two opaque tensors and forward*() picks it
    up.  The setting is recorded PER GRAPH at preprocess time: flipping the default afterwards changes nothing for
    this graph, and a graph preprocessed with the default off carries no plan."""
    rp, ci = GRAPHS["sbm_1024"]
    n = 1024
    H.set_dense(True)
    try:
        pre = prep(H, rp, ci, "all_tc")
        assert pre[4].device.type == "cpu" and int(pre[4][1]) == 8 and int(pre[4][6]) == 1 and pre[5].numel() > 16
        g = torch.Generator().manual_seed(7)
        x, w = torch.randn(n, 64, generator=g), torch.randn(64, 32, generator=g)
        tf32 = oracle.spmm(rp, ci, oracle.tf32_round(x.numpy()), precision=1)
        out = H.forward(x.cuda(), dev(rp), dev(ci), *pre)[0]
        assert rel_fro(out.cpu().numpy(), tf32) <= 2e-5
        o2, z = H.forward_fixed32_fused(x.cuda(), dev(rp), dev(ci), *pre, w.cuda())
        assert rel_fro(z.cpu().numpy(), tf32) <= 2e-5
        assert rel_fro(o2.cpu().numpy(), oracle.gemm(tf32, w.numpy(), tf32=True)) <= 1e-4
        # shipped selector: no tensor-core windows, hence no plan
        pre0 = prep(H, rp, ci, "shipped")
        assert int(pre0[4][1]) == 0 and int(pre0[4][7]) == 0        # no dense groups, no label-1 windows
    finally:
        H.set_dense(False)
    out = H.forward(x.cuda(), dev(rp), dev(ci), *pre)[0]      # this graph keeps the plan it was preprocessed with
    assert rel_fro(out.cpu().numpy(), tf32) <= 2e-5
    pre1 = prep(H, rp, ci, "all_tc")                            # default off again: no plan, per-window tensor-core path
    assert int(pre1[4][1]) == 0 and int(pre1[4][6]) == 0
    out = H.forward(x.cuda(), dev(rp), dev(ci), *pre1)[0]
    assert rel_fro(out.cpu().numpy(), tf32) <= 2e-5
