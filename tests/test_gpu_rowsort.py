"""Row-sorted copy of the CSR for the balanced kernel (csrc/rowsort.cu, hcspmm_row_sort): the permutation is the stable
sort of the rows by length class (bit-exact against numpy), every row keeps its entries in order, and the SpMM through
the sorted copy equals the oracle -- with labels, accumulate, BF16 and empty rows -- and equals the unsorted path."""
import numpy as np
import pytest
import torch

import oracle
from helpers import rel_fro, small_graphs

pytestmark = pytest.mark.gpu


def _classes(deg):
    d = np.maximum(deg, 1).astype(np.int64)
    return np.where(d <= 1, 31, 31 - np.floor(np.log2(d)).astype(np.int64))          # __clz of a 32-bit int


@pytest.mark.parametrize("name", ["rmat_hub_4096", "rmat_1000", "holes_777", "uniform_777", "empty_48", "single_row",
                                  "dense_2048", "ring3_256"])
def test_row_sort_is_the_stable_class_sort(name):
    from hcspmm import capi
    rp, ci = small_graphs()[name]
    n = rp.size - 1
    dev = torch.device("cuda", 0)
    row_id, rp_s, ci_s = capi.row_sort_csr(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev))
    torch.cuda.synchronize()
    deg = np.diff(rp)
    want_id = np.argsort(_classes(deg), kind="stable")
    assert np.array_equal(row_id.cpu().numpy(), want_id)
    want_rp = np.zeros(n + 1, np.int64)
    want_rp[1:] = np.cumsum(deg[want_id])
    assert np.array_equal(rp_s.cpu().numpy(), want_rp)
    want_ci = np.concatenate([ci[rp[r]:rp[r + 1]] for r in want_id]) if ci.size else ci
    assert np.array_equal(ci_s.cpu().numpy(), want_ci)


@pytest.mark.parametrize("dim", [128, 64, 100, 256])
@pytest.mark.parametrize("name", ["rmat_hub_4096", "holes_777", "rmat_1000"])
def test_spmm_through_the_sorted_copy_matches_oracle(name, dim):
    from hcspmm import capi
    rp, ci = small_graphs()[name]
    n = rp.size - 1
    dev = torch.device("cuda", 0)
    d_rp, d_ci = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
    x = np.random.default_rng(dim).standard_normal((n, dim)).astype(np.float32)
    want = oracle.spmm(rp, ci, x, precision=1)
    d_x = torch.from_numpy(x).to(dev)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "shipped")
    old = capi.set_tuning("balance", 2)                      # the balanced kernel whatever the mean degree
    try:
        plain = capi.spmm_aux(d_x, d_rp, d_ci, bp, etc, etr, ht, capi.GraphAux(d_rp, d_ci, ht, row_sort=False))
        aux = capi.GraphAux(d_rp, d_ci, ht, row_sort=True)
        assert aux.sorted is not None
        got = capi.spmm_aux(d_x, d_rp, d_ci, bp, etc, etr, ht, aux)
        assert rel_fro(got.cpu().numpy(), want) <= 1e-5
        assert rel_fro(got.cpu().numpy(), plain.cpu().numpy()) <= 1e-6
        # accumulate into a pre-filled Y, and the BF16 precision mode
        y0 = torch.randn(n, dim, device=dev)
        acc = capi.spmm_aux(d_x, d_rp, d_ci, bp, etc, etr, ht, aux, out=y0.clone(), accumulate=True)
        assert rel_fro((acc - y0).cpu().numpy(), want) <= 1e-4
        if dim % 8 == 0:
            b16 = capi.spmm_aux(d_x, d_rp, d_ci, bp, etc, etr, ht, aux, precision="bf16")
            assert rel_fro(b16.cpu().numpy(), want) <= 1e-2
    finally:
        capi.set_tuning("balance", old)


def test_sorted_copy_respects_window_labels():
    """Windows labelled tensor-core are computed by the per-window kernel on the ORIGINAL CSR; the balanced kernel, walking
    the sorted copy, must skip exactly their rows (labels are looked up through row_id)."""
    from hcspmm import capi
    rp, ci = small_graphs()["rmat_hub_4096"]
    n = rp.size - 1
    dev = torch.device("cuda", 0)
    d_rp, d_ci = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
    x = np.random.default_rng(1).standard_normal((n, 64)).astype(np.float32)
    want = oracle.spmm(rp, ci, x, precision=1)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "intended")
    ht = ht.clone()
    ht[::3] = 1                                              # a third of the windows on the mma.sync path
    aux = capi.GraphAux(d_rp, d_ci, ht, row_sort=True)
    got = capi.spmm_aux(torch.from_numpy(x).to(dev), d_rp, d_ci, bp, etc, etr, ht, aux, precision="tf32")
    assert rel_fro(got.cpu().numpy(), want) <= 1e-3


def test_module_preprocess_carries_the_sorted_copy():
    import HCSPMM
    from hcspmm import graphs
    dev = torch.device("cuda", 0)
    rp, ci = graphs.rmat(20_000, 500_000, seed=4)            # mean row 25: inside the rule (8 <= mean < 64)
    d_rp, d_ci = rp.to(dev), ci.to(dev)
    n = rp.numel() - 1
    x = torch.randn(n, 128, device=dev)
    HCSPMM.set_classifier("shipped")
    pre = HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16)
    assert int(pre[4][15]) > 0, "preprocess did not emit the row-sorted copy"
    y = HCSPMM.forward(x, d_rp, d_ci, *pre)[0]
    old = HCSPMM.set_row_sort(False)
    try:
        pre0 = HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16)
    finally:
        HCSPMM.set_row_sort(old)
    assert int(pre0[4][15]) == 0
    y0 = HCSPMM.forward(x, d_rp, d_ci, *pre0)[0]
    assert float((y - y0).norm() / y0.norm()) <= 1e-6
    want = oracle.spmm(rp.numpy(), ci.numpy(), x.cpu().numpy(), precision=1)
    assert rel_fro(y.cpu().numpy(), want) <= 1e-5
