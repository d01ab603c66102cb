"""CPU tests of the oracle itself (no GPU): the C restatement against independent numpy /
torch.sparse computations, the classifier's decision boundary, and the LOA
restatement against the unmodified reference LOI.cpp when oracle/_ref/libloi_ref.so exists."""
import numpy as np
import pytest
import torch

import oracle
from helpers import rel_fro, small_graphs, torch_sparse_ref

GRAPHS = small_graphs()


def numpy_preprocess(rowptr, colidx):
    """Independent restatement with numpy set operations (no sort/dedup/binary-search loops)."""
    n = rowptr.size - 1
    w = oracle.num_windows(n)
    bp = np.zeros(w, np.int32)
    etc = np.zeros(colidx.size, np.int32)
    etr = np.repeat(np.arange(n, dtype=np.int32), np.diff(rowptr))
    uniq = np.zeros(w, np.int64)
    for i in range(w):
        e0, e1 = rowptr[i * 16], rowptr[min(i * 16 + 16, n)]
        if e1 == e0:
            continue
        u, inv = np.unique(colidx[e0:e1], return_inverse=True)
        etc[e0:e1] = inv
        uniq[i] = u.size
        bp[i] = (u.size + 7) // 8
    return bp, etc, etr, uniq


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_preprocess_matches_numpy(name):
    rp, ci = GRAPHS[name]
    bp, etc, etr, ht = oracle.preprocess(ci, rp, oracle.MODE_SHIPPED)
    bp2, etc2, etr2, _ = numpy_preprocess(rp, ci)
    assert np.array_equal(bp, bp2)
    assert np.array_equal(etc, etc2)
    assert np.array_equal(etr, etr2)
    assert not ht.any()  # shipped selector (hybrid_all_kernel.cu:262): label 0 unless score == 0.0


def test_classifier_boundary():
    """Decision boundary of the intended rule (hybrid_all_kernel.cu:261) as SURVEY.md 2.3 derives it:
    <= 17 distinct columns always TC; 24 columns need >= 83 edges; 25 need >= 126; 33 need >= 312;
    34+ columns are always CUDA."""
    lab = oracle.lib().hcspmm_oracle_label

    def tc(u, e):
        return lab(u - 1, e, (u - 1 + 8) // 8, oracle.MODE_INTENDED)

    for u in range(1, 18):
        assert tc(u, u) == 1
    assert tc(24, 82) == 0 and tc(24, 83) == 1
    assert tc(25, 125) == 0 and tc(25, 126) == 1
    assert tc(33, 311) == 0 and tc(33, 312) == 1
    for u in (34, 35, 64, 500):
        assert tc(u, 16 * u) == 0
    # shipped rule: 0 for every realistic window
    assert lab(23, 83, 3, oracle.MODE_SHIPPED) == 0


def test_classifier_score_is_double_fma():
    """score = fma((double)(float)size, w1, (double)((float)E / (float)(num*128)) * -w2) + -b"""
    import math
    size, e, num = 23, 83, 3
    d = float(np.float32(e) / np.float32(num * 128))
    want = math.fma(float(np.float32(size)), 0.19854024, d * -6.578043) + -3.14922857 \
        if hasattr(math, "fma") else None
    got = oracle.lib().hcspmm_oracle_score(size, e, num)
    if want is not None:
        assert got == want
    assert abs(got - (size * 0.19854024 - d * 6.578043 - 3.14922857)) < 1e-12


@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_spmm_fp32_matches_torch_sparse(name):
    rp, ci = GRAPHS[name]
    n = rp.size - 1
    xr = int(ci.max()) + 1 if ci.size else n
    xr = max(xr, n)
    x = np.random.default_rng(0).standard_normal((xr, 24)).astype(np.float32)
    y = oracle.spmm(rp, ci, x, precision=1)
    ref = torch_sparse_ref(rp, ci, x)
    assert rel_fro(y, ref) <= 1e-6


def test_spmm_tf32_rounding_and_accumulate():
    rp, ci = GRAPHS["band2_320"]
    x = np.random.default_rng(1).standard_normal((320, 16)).astype(np.float32)
    ht = np.ones(20, np.int32)
    y_tc = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0)
    y_ref = oracle.spmm(rp, ci, oracle.tf32_round(x), precision=1)
    assert np.array_equal(y_tc, y_ref)
    assert 0 < rel_fro(y_tc, oracle.spmm(rp, ci, x, precision=1)) <= 1e-3
    base = np.ones((320, 16), np.float32)
    y_acc = oracle.spmm(rp, ci, x, precision=1, y_init=base)
    assert np.allclose(y_acc, base + oracle.spmm(rp, ci, x, precision=1), atol=1e-6)


def test_tf32_round_ties_away():
    one = np.float32(1.0)
    ulp = np.float32(2.0 ** -10)       # TF32 spacing at 1.0
    half = np.float32(1.0 + 2.0 ** -11)
    f = oracle.lib().hcspmm_oracle_tf32
    assert f(float(half)) == float(one + ulp)          # tie -> away from zero
    assert f(float(-half)) == float(-(one + ulp))
    assert f(float(np.float32(1.0 + 2.0 ** -12))) == 1.0
    arr = np.array([half, -half, 3.14159], np.float32)
    assert np.array_equal(oracle.tf32_round(arr), np.array([f(float(v)) for v in arr], np.float32))


def test_gemm_oracle():
    rng = np.random.default_rng(2)
    z = rng.standard_normal((50, 40)).astype(np.float32)
    w = rng.standard_normal((40, 24)).astype(np.float32)
    assert rel_fro(oracle.gemm(z, w), z.astype(np.float64) @ w.astype(np.float64)) < 1e-6
    assert rel_fro(oracle.gemm(z, w, tf32=True), z.astype(np.float64) @ w.astype(np.float64)) < 2e-3


# ---- LOA ---------------------------------------------------------------------------------
LOA_GRAPHS = ["ring3_256", "band2_320", "rmat_1000", "sbm_1024", "uniform_777", "holes_777", "empty_48"]


@pytest.mark.parametrize("name", LOA_GRAPHS)
def test_loa_is_a_permutation(name):
    rp, ci = GRAPHS[name]
    n = rp.size - 1
    perm, sizes, n_full = oracle.loa(rp, ci)
    assert np.array_equal(np.sort(perm), np.arange(n))
    assert sizes.sum() == np.count_nonzero(np.diff(rp))  # every vertex with an edge is in a block
    assert (sizes <= 16).all() and (sizes >= 1).all()
    assert n_full == np.count_nonzero(sizes == 16)


@pytest.mark.skipif(not oracle.have_loi_ref(), reason="oracle/_ref/libloi_ref.so not built (no reference here)")
@pytest.mark.parametrize("name", LOA_GRAPHS)
def test_loa_matches_reference_loi(name):
    """Pins the LOA restatement against the UNMODIFIED reference LOI.cpp (reorder_plus_new_direct)."""
    rp, ci = GRAPHS[name]
    perm, sizes, n_full = oracle.loa(rp, ci)
    rperm, rsizes, rfull = oracle.loa_reference(rp, ci)
    assert np.array_equal(sizes, rsizes)
    assert n_full == rfull
    assert np.array_equal(perm, rperm)
