"""GPU parity tests (B200): the CUDA path, called through the C ABI (hcspmm.capi -> libhcspmm.so),
against the CPU oracle on the same seeded inputs.

Bars: integer / index outputs bit-exact; FP32 (CUDA-core) windows <= 1e-5 normwise; TF32
(tensor-core) windows <= 1e-3 normwise against the FP32 result (north star), BF16 n/a here."""
import numpy as np
import pytest
import torch

import oracle
from helpers import rel_fro, small_graphs, torch_sparse_ref

pytestmark = pytest.mark.gpu

GRAPHS = small_graphs()
TOL_FP32 = 1e-5
TOL_TF32 = 1e-3


@pytest.fixture(scope="module")
def capi():
    from hcspmm import capi as c
    c.lib()
    return c


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def xmat(rows, dim, seed=0):
    return np.random.default_rng(seed).standard_normal((rows, dim)).astype(np.float32)


def x_rows_for(rp, ci):
    n = rp.size - 1
    return max(n, int(ci.max()) + 1 if ci.size else n)


# ---- A1-A5 preprocessing: bit-exact ---------------------------------------------------------
@pytest.mark.parametrize("mode", ["shipped", "intended"])
@pytest.mark.parametrize("name", sorted(GRAPHS))
def test_preprocess_bit_exact(capi, name, mode):
    rp, ci = GRAPHS[name]
    want = oracle.preprocess(ci, rp, capi.CLASSIFIERS[mode])
    got = capi.preprocess(dev(ci), dev(rp), mode)
    for g, w, what in zip(got, want, ("blockPartition", "edgeToColumn", "edgeToRow", "hybrid_type")):
        assert np.array_equal(g.cpu().numpy(), w), f"{what} differs on {name}/{mode}"


def test_preprocess_forced_modes(capi):
    rp, ci = GRAPHS["holes_777"]
    nonempty = np.array([rp[min(16 * w + 16, rp.size - 1)] > rp[16 * w] for w in range(oracle.num_windows(rp.size - 1))])
    ht = capi.preprocess(dev(ci), dev(rp), "all_tc")[3].cpu().numpy()
    assert np.array_equal(ht != 0, nonempty)        # empty windows stay 0
    assert not capi.preprocess(dev(ci), dev(rp), "all_cuda")[3].any()


def test_preprocess_rejects_bad_window_count(capi):
    rp, ci = GRAPHS["ring3_256"]
    d_ci, d_rp = dev(ci), dev(rp)
    buf = torch.zeros(1024, dtype=torch.int32, device="cuda")
    rc = capi.lib().hcspmm_preprocess(d_ci.data_ptr(), d_rp.data_ptr(), 256, ci.size, 15, 0, buf.data_ptr(),
                                      buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4096, None)
    assert rc == -1 and b"n_windows" in capi.lib().hcspmm_last_error()


# ---- A6/A7 SpMM, CUDA-core path ----------------------------------------------------------------
DIMS = [1, 3, 22, 32, 47, 64, 96, 100, 128, 200, 256, 512, 640]


@pytest.mark.parametrize("dim", DIMS)
@pytest.mark.parametrize("name", ["ring3_256", "rmat_1000", "holes_777", "empty_48", "single_row", "sbm_1024"])
def test_spmm_fp32(capi, name, dim):
    rp, ci = GRAPHS[name]
    x = xmat(x_rows_for(rp, ci), dim, seed=dim)
    want = oracle.spmm(rp, ci, x, precision=1)
    got = capi.spmm(dev(x), dev(rp), dev(ci), precision="fp32").cpu().numpy()
    assert got.shape == want.shape
    assert rel_fro(got, want) <= TOL_FP32
    if ci.size:
        assert rel_fro(got, torch_sparse_ref(rp, ci, x)) <= TOL_FP32


@pytest.mark.parametrize("long_row", [32, 100, 1024])
@pytest.mark.parametrize("dim", [32, 64, 128, 256, 512])
def test_spmm_long_rows_split_over_warps(capi, dim, long_row):
    rp, ci = GRAPHS["rmat_hub_4096"]
    x = xmat(4096, dim, seed=3)
    want = oracle.spmm(rp, ci, x, precision=1)
    old = capi.set_tuning("long_row", long_row)
    try:
        got = capi.spmm(dev(x), dev(rp), dev(ci), precision="fp32").cpu().numpy()
    finally:
        capi.set_tuning("long_row", old)
    assert rel_fro(got, want) <= TOL_FP32


@pytest.mark.parametrize("slab", [32, 64, 96, 128])
def test_spmm_feature_slabs(capi, slab):
    rp, ci = GRAPHS["rmat_1000"]
    ht = np.zeros(oracle.num_windows(1000), np.int32)
    ht[::2] = 1
    bp, etc, etr, _ = oracle.preprocess(ci, rp, 0)
    for dim in (256, 200, 520):
        x = xmat(1000, dim, seed=slab)
        want = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0)
        old = capi.set_tuning("slab", slab)
        try:
            got = capi.spmm(dev(x), dev(rp), dev(ci), dev(bp), dev(etc), dev(etr), dev(ht)).cpu().numpy()
        finally:
            capi.set_tuning("slab", old)
        assert rel_fro(got, want) <= 2e-5


@pytest.mark.parametrize("wpc", [1, 2, 8])
@pytest.mark.parametrize("short_row", [0, 8, 100000])
def test_spmm_windows_per_cta_and_group_rows(capi, wpc, short_row):
    """Low-degree machinery: several windows per CTA, one lane group per short row; mixed labels."""
    for name in ("ring3_256", "rmat_1000", "holes_777", "sbm_1024"):
        rp, ci = GRAPHS[name]
        n = rp.size - 1
        bp, etc, etr, _ = oracle.preprocess(ci, rp, 0)
        ht = np.zeros(oracle.num_windows(n), np.int32)
        ht[2::5] = 1
        for dim in (8, 32, 64, 128, 256):
            x = xmat(n, dim, seed=dim + wpc)
            want = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0)
            o1, o2 = capi.set_tuning("wpc", wpc), capi.set_tuning("short_row", short_row)
            try:
                got = capi.spmm(dev(x), dev(rp), dev(ci), dev(bp), dev(etc), dev(etr), dev(ht)).cpu().numpy()
                acc = dev(np.ones_like(want))
                capi.spmm(dev(x), dev(rp), dev(ci), dev(bp), dev(etc), dev(etr), dev(ht), out=acc, accumulate=True)
            finally:
                capi.set_tuning("wpc", o1), capi.set_tuning("short_row", o2)
            assert rel_fro(got, want) <= 2e-5, (name, dim)
            assert rel_fro(acc.cpu().numpy(), want + 1.0) <= 2e-5, (name, dim)


@pytest.mark.parametrize("chunk", [64, 100, 777, 4096])
@pytest.mark.parametrize("warp_split", [-1, 0, 1, 64])
def test_spmm_balanced_items(capi, chunk, warp_split):
    """Merge-path balanced CUDA-core kernel: rows cut by item boundaries (hubs), empty rows / windows,
    mixed labels (tensor-core windows skipped there and computed by the per-window kernel), accumulate,
    BF16-stored X.  warp_split -1 = the per-window kernel (balance off), which must agree; 0 = warp-per-row /
    CTA-per-long-row phases inside every item; 1 / 64 = equal entry runs per warp in items of long enough rows."""
    for name in ("rmat_hub_4096", "holes_777", "rmat_1000", "empty_48", "single_row", "rect_72x100000", "dense_2048"):
        rp, ci = GRAPHS[name]
        n = rp.size - 1
        bp, etc, etr, _ = oracle.preprocess(ci, rp, 0)
        ht = np.zeros(oracle.num_windows(n), np.int32)
        ht[1::4] = 1
        for dim in (32, 100, 256, 520):
            x = xmat(x_rows_for(rp, ci), dim, seed=dim + chunk)
            y0 = xmat(n, dim, seed=7)
            want32 = oracle.spmm(rp, ci, x, precision=1)
            want_h = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0 if dim % 8 == 0 else 1)
            o1, o2 = capi.set_tuning("balance", 0 if warp_split < 0 else 2), capi.set_tuning("chunk", chunk)
            o3 = capi.set_tuning("warp_split", max(warp_split, 0))
            try:
                got32 = capi.spmm(dev(x), dev(rp), dev(ci), precision="fp32").cpu().numpy()
                got_h = capi.spmm(dev(x), dev(rp), dev(ci), dev(bp), dev(etc), dev(etr), dev(ht)).cpu().numpy()
                acc = dev(y0)
                capi.spmm(dev(x), dev(rp), dev(ci), precision="fp32", out=acc, accumulate=True)
                got16 = capi.spmm(dev(x), dev(rp), dev(ci), precision="bf16").cpu().numpy()
            finally:
                capi.set_tuning("balance", o1), capi.set_tuning("chunk", o2), capi.set_tuning("warp_split", o3)
            assert rel_fro(got32, want32) <= TOL_FP32, (name, dim)
            assert rel_fro(got_h, want_h) <= 2e-5, (name, dim)
            assert rel_fro(acc.cpu().numpy(), want32 + y0) <= 2e-5, (name, dim)
            assert rel_fro(got16, want32) <= 1e-2, (name, dim)


def test_spmm_balanced_is_deterministic(capi):
    rp, ci = GRAPHS["rmat_hub_4096"]
    x = dev(xmat(4096, 128, seed=2))
    old, ob = capi.set_tuning("chunk", 256), capi.set_tuning("balance", 2)
    try:
        a = capi.spmm(x, dev(rp), dev(ci), precision="fp32")
        for _ in range(3):
            assert torch.equal(a, capi.spmm(x, dev(rp), dev(ci), precision="fp32"))
    finally:
        capi.set_tuning("chunk", old), capi.set_tuning("balance", ob)


def test_spmm_strided_and_unaligned(capi):
    rp, ci = GRAPHS["rmat_1000"]
    big = dev(xmat(1000, 80, seed=9))
    want_full = oracle.spmm(rp, ci, big.cpu().numpy(), precision=1)
    # row-strided view, 16-byte aligned start: vector kernel with ldx != dim
    v = big[:, 16:48]
    got = capi.spmm(v, dev(rp), dev(ci), precision="fp32").cpu().numpy()
    assert rel_fro(got, want_full[:, 16:48]) <= TOL_FP32
    # 4-byte aligned start: scalar kernel
    v = big[:, 3:35]
    got = capi.spmm(v, dev(rp), dev(ci), precision="fp32").cpu().numpy()
    assert rel_fro(got, want_full[:, 3:35]) <= TOL_FP32


def test_spmm_accumulate(capi):
    rp, ci = GRAPHS["rmat_1000"]
    bp, etc, etr, _ = oracle.preprocess(ci, rp, 0)
    ht = np.zeros(oracle.num_windows(1000), np.int32)
    ht[1::3] = 1
    for dim in (32, 100, 47):
        x = xmat(1000, dim, seed=4)
        y0 = xmat(1000, dim, seed=5)
        # dim % 8 != 0: tensor-core windows fall back to the exact FP32 path
        want = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0 if dim % 8 == 0 else 1, y_init=y0)
        out = dev(y0)
        capi.spmm(dev(x), dev(rp), dev(ci), dev(bp), dev(etc), dev(etr), dev(ht), out=out, accumulate=True)
        assert rel_fro(out.cpu().numpy(), want) <= 2e-5


def test_spmm_rectangular_shard(capi):
    """Row-partitioned shard: 72 output rows, 100000 rows of X, global column ids."""
    rp, ci = GRAPHS["rect_72x100000"]
    for classifier in ("all_cuda", "all_tc"):
        bp, etc, etr, ht = capi.preprocess(dev(ci), dev(rp), classifier)
        x = xmat(100000, 64, seed=6)
        got = capi.spmm(dev(x), dev(rp), dev(ci), bp, etc, etr, ht).cpu().numpy()
        assert got.shape == (72, 64)
        want = oracle.spmm(rp, ci, x, precision=1)
        assert rel_fro(got, want) <= (TOL_FP32 if classifier == "all_cuda" else TOL_TF32)
    # column ids beyond x_rows contribute zero (reference TC path does the same, :1093)
    xs = x[:50000]
    keep = ci < 50000
    deg = np.array([keep[rp[r]:rp[r + 1]].sum() for r in range(rp.size - 1)])
    rp_f = np.zeros_like(rp)
    rp_f[1:] = np.cumsum(deg)
    want = torch_sparse_ref(rp_f, ci[keep], xs)
    assert rel_fro(oracle.spmm(rp, ci, xs, precision=1), want) <= 1e-6
    got = capi.spmm(dev(xs), dev(rp), dev(ci), precision="fp32").cpu().numpy()
    assert rel_fro(got, want) <= TOL_FP32


# ---- tensor-core path -----------------------------------------------------------------------
@pytest.mark.parametrize("dim", [8, 16, 32, 48, 64, 128, 256, 512])
@pytest.mark.parametrize("name", ["band2_320", "sbm_1024", "rmat_1000", "holes_777", "dense_2048"])
def test_spmm_tensor_core_windows(capi, name, dim):
    rp, ci = GRAPHS[name]
    bp, etc, etr, ht = capi.preprocess(dev(ci), dev(rp), "all_tc")
    x = xmat(x_rows_for(rp, ci), dim, seed=dim + 1)
    fp32 = oracle.spmm(rp, ci, x, precision=1)
    tf32 = oracle.spmm(rp, ci, x, hybrid_type=ht.cpu().numpy(), precision=0)
    got = capi.spmm(dev(x), dev(rp), dev(ci), bp, etc, etr, ht, precision="tf32").cpu().numpy()
    assert rel_fro(got, fp32) <= TOL_TF32            # the north star's bar
    assert rel_fro(got, tf32) <= 2e-5                # same rounding as the oracle's TF32 restatement
    got2 = capi.spmm(dev(x), dev(rp), dev(ci), bp, etc, etr, ht, precision="tf32x2").cpu().numpy()
    assert rel_fro(got2, fp32) <= TOL_FP32           # split-TF32 recovers FP32 accuracy


def test_spmm_mixed_labels_intended_selector(capi):
    """A graph whose windows straddle the intended selector's boundary: both paths in one launch."""
    rng = np.random.default_rng(8)
    n = 640
    rows = []
    for r in range(n):
        w = r // 16
        span = 12 + (w % 5) * 9           # 12..48 distinct candidate columns per window
        k = int(rng.integers(1, min(span, 14)))
        rows.append(np.sort(rng.choice(span, size=k, replace=False) + (w * 7) % (n - 64)).astype(np.int32))
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    ci = np.concatenate(rows)
    want_pre = oracle.preprocess(ci, rp, oracle.MODE_INTENDED)
    got_pre = capi.preprocess(dev(ci), dev(rp), "intended")
    for g, w in zip(got_pre, want_pre):
        assert np.array_equal(g.cpu().numpy(), w)
    ht = want_pre[3]
    assert 0 < ht.sum() < ht.size, "fixture must mix CUDA-core and tensor-core windows"
    x = xmat(n, 32, seed=2)
    got = capi.spmm(dev(x), dev(rp), dev(ci), *got_pre).cpu().numpy()
    want = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0)
    assert rel_fro(got, want) <= 2e-5
    assert rel_fro(got, oracle.spmm(rp, ci, x, precision=1)) <= TOL_TF32


# ---- A8 fused aggregation + update ----------------------------------------------------------
@pytest.mark.parametrize("dim,hidden", [(32, 32), (22, 32), (100, 47), (128, 128), (256, 64), (64, 256), (5, 3)])
def test_spmm_gemm(capi, dim, hidden):
    rp, ci = GRAPHS["rmat_1000"]
    pre = capi.preprocess(dev(ci), dev(rp), "shipped")
    x = xmat(1000, dim, seed=7)
    w = xmat(dim, hidden, seed=8)
    out, z = capi.spmm_gemm(dev(x), dev(rp), dev(ci), *pre, dev(w))
    z_want = oracle.spmm(rp, ci, x, precision=1)
    assert rel_fro(z.cpu().numpy(), z_want) <= TOL_FP32
    assert rel_fro(out.cpu().numpy(), oracle.gemm(z_want, w, tf32=True)) <= 1e-4
    assert rel_fro(out.cpu().numpy(), z_want.astype(np.float64) @ w.astype(np.float64)) <= 2e-3


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (127, 33, 65), (300, 16, 64), (1000, 128, 128), (513, 100, 47)])
def test_gemm_tf32(capi, m, k, n):
    a, b = xmat(m, k, 1), xmat(k, n, 2)
    got = capi.gemm_tf32(dev(a), dev(b)).cpu().numpy()
    assert rel_fro(got, oracle.gemm(a, b, tf32=True)) <= 1e-4


# ---- host-buffer path -------------------------------------------------------------------------
def test_host_graph_roundtrip(capi):
    rp, ci = GRAPHS["rmat_1000"]
    g = capi.HostGraph(torch.from_numpy(rp), torch.from_numpy(ci), classifier="intended")
    want_pre = oracle.preprocess(ci, rp, oracle.MODE_INTENDED)
    for a, b in zip(g.preprocess_arrays(), want_pre):
        assert np.array_equal(a.numpy(), b)
    x = torch.from_numpy(xmat(1000, 96, seed=3)).pin_memory()
    y = g.spmm(x)
    assert rel_fro(y.numpy(), oracle.spmm(rp, ci, x.numpy(), hybrid_type=want_pre[3], precision=0)) <= 2e-5
    g.close()


# ---- tcgen05 / TMEM Update GEMM ---------------------------------------------------------------
@pytest.mark.parametrize("m,k,n", [(128, 32, 32), (128, 64, 256), (1000, 128, 128), (513, 100, 48), (4096, 256, 256),
                                   (300, 36, 16), (2000, 128, 320), (77, 8, 4)])
def test_gemm_tcgen05(capi, m, k, n):
    a, b = xmat(m, k, 1), xmat(k, n, 2)
    old = capi.set_tuning("umma_gemm", 1)
    try:
        got = capi.gemm_tf32(dev(a), dev(b)).cpu().numpy()
        assert capi.lib().hcspmm_debug_umma_error() == 0, "tcgen05 kernel reported a barrier timeout"
    finally:
        capi.set_tuning("umma_gemm", old)
    assert rel_fro(got, oracle.gemm(a, b, tf32=True)) <= 1e-4


# ---- per-graph products (hcspmm_aux_t) and the BF16-stored operand ------------------------------
@pytest.mark.parametrize("name", ["rmat_1000", "rmat_hub_4096", "dense_2048", "holes_777", "rect_72x100000"])
@pytest.mark.parametrize("dim", [64, 256])
def test_spmm_aux_is_the_same_computation(capi, name, dim):
    """Precomputed merge-path split points (base chunk 4096, read with stride 1 at dim 256 and stride 2 at dim 64) and
    a caller workspace change nothing: bit-identical to hcspmm_spmm, and within FP32 tolerance of the oracle."""
    rp, ci = GRAPHS[name]
    x_rows = 100000 if name.startswith("rect") else rp.size - 1
    x = xmat(x_rows, dim, seed=5)
    pre = capi.preprocess(dev(ci), dev(rp), "shipped")
    old = capi.set_tuning("balance", 2)
    try:
        plain = capi.spmm(dev(x), dev(rp), dev(ci), *pre)
        aux = capi.GraphAux(dev(rp), dev(ci), pre[3], row_sort=False)      # the sorted copy: tests/test_gpu_rowsort.py
        assert aux.n_tc == int((pre[3] == 1).sum())
        got = capi.spmm_aux(dev(x), dev(rp), dev(ci), *pre, aux)
        again = capi.spmm_aux(dev(x), dev(rp), dev(ci), *pre, aux)
    finally:
        capi.set_tuning("balance", old)
    assert torch.equal(got, plain) and torch.equal(again, plain)
    assert rel_fro(got.cpu().numpy(), oracle.spmm(rp, ci, x, precision=1)) <= TOL_FP32


def test_f32_to_bf16_is_round_to_nearest_even(capi):
    x = torch.randn(1000, 96, generator=torch.Generator().manual_seed(3))
    x[0, :4] = torch.tensor([float("inf"), -float("inf"), 0.0, -0.0])
    got = capi.f32_to_bf16(x.cuda())
    assert torch.equal(got.cpu().view(torch.int16), x.to(torch.bfloat16).view(torch.int16))
    wide = torch.zeros(1000, 128, dtype=torch.bfloat16, device="cuda")     # strided destination (operand segment)
    capi.f32_to_bf16(x.cuda(), wide[:, 16:112])
    assert torch.equal(wide[:, 16:112].cpu().view(torch.int16), x.to(torch.bfloat16).view(torch.int16))
    assert not wide[:, :16].any() and not wide[:, 112:].any()


@pytest.mark.parametrize("name", ["rmat_hub_4096", "uniform_777", "rect_72x100000"])
def test_spmm_bf16_stored_operand(capi, name):
    """A bfloat16-stored X (the multi-GPU exchange operand) gives exactly the BF16 precision mode's result."""
    rp, ci = GRAPHS[name]
    x_rows = 100000 if name.startswith("rect") else rp.size - 1
    x = xmat(x_rows, 128, seed=6)
    pre = capi.preprocess(dev(ci), dev(rp), "shipped")
    want = capi.spmm(dev(x), dev(rp), dev(ci), *pre, precision="bf16")
    xb = capi.f32_to_bf16(dev(x))
    aux = capi.GraphAux(dev(rp), dev(ci), pre[3], row_sort=False)
    got = capi.spmm_aux(xb, dev(rp), dev(ci), *pre, aux, precision="bf16_stored")
    assert torch.equal(got, want)
    assert rel_fro(got.cpu().numpy(), oracle.spmm(rp, ci, x, precision=1)) <= 1e-2


# ---- TMA + tcgen05 persistent Update GEMM (csrc/update_gemm.cu) ---------------------------------
@pytest.mark.parametrize("m,k,n", [(128, 32, 32), (128, 64, 256), (1000, 128, 128), (513, 100, 48), (4096, 256, 256),
                                   (300, 36, 16), (2000, 128, 320), (77, 8, 4), (40000, 128, 128), (19000, 100, 128),
                                   (40000, 47, 128), (40000, 128, 47), (30000, 47, 47)])   # odd widths: padded copies
@pytest.mark.parametrize("rounders", [1, 0])
def test_gemm_tma(capi, m, k, n, rounders):
    """rounders = 1: cvt.rna in shared memory (the reference's rounding, oracle-tight); 0: the tensor map's
    TF32 element type (hardware conversion on load, tolerance of one TF32 ulp per operand)."""
    a, b = xmat(m, k, 1), xmat(k, n, 2)
    old = capi.set_tuning("umma_gemm", 2)
    old_r = capi.set_tuning("gemm_round", rounders)
    try:
        got = capi.gemm_tf32(dev(a), dev(b)).cpu().numpy()
        assert capi.lib().hcspmm_debug_umma_error() == 0, "tcgen05 kernel reported a barrier timeout"
    finally:
        capi.set_tuning("umma_gemm", old)
        capi.set_tuning("gemm_round", old_r)
    assert rel_fro(got, oracle.gemm(a, b, tf32=True)) <= (1e-4 if rounders else 2e-3)


def test_gemm_tma_strided_views(capi):
    """Z / out as column blocks of wider matrices (lda, ldo > width) and a strided W."""
    a_full, b_full = xmat(3000, 160, 3), xmat(128, 200, 4)
    a, b = a_full[:, 16:144], b_full[:, 8:136]
    d_a, d_b = dev(a_full)[:, 16:144], dev(b_full)[:, 8:136]
    out_full = torch.zeros(3000, 256, device="cuda")
    old = capi.set_tuning("umma_gemm", 2)
    try:
        with torch.cuda.device(0):
            capi._check(capi.lib().hcspmm_gemm_tf32(d_a.data_ptr(), 160, d_b.data_ptr(), 200, 3000, 128, 128,
                                                    out_full[:, 64:192].data_ptr(), 256,
                                                    torch.cuda.current_stream().cuda_stream), "hcspmm_gemm_tf32")
        torch.cuda.synchronize()
    finally:
        capi.set_tuning("umma_gemm", old)
    got = out_full.cpu().numpy()
    assert rel_fro(got[:, 64:192], oracle.gemm(np.ascontiguousarray(a), np.ascontiguousarray(b), tf32=True)) <= 1e-4
    assert not got[:, :64].any() and not got[:, 192:].any(), "the TMA store wrote outside its column block"


def test_spmm_odd_width_large_uses_padded_copies(capi):
    """dim = 47 on a graph large enough for the pad path (n_rows * dim >= 2^20), with a hub row."""
    from hcspmm import graphs as G
    rp, ci = G.rmat(40000, 600000, seed=9)
    rp, ci = rp.numpy(), ci.numpy()
    for dim in (47, 30):
        x = xmat(40000, dim, seed=dim)
        want = oracle.spmm(rp, ci, x, precision=1)
        got = capi.spmm(dev(x), dev(rp), dev(ci), precision="fp32").cpu().numpy()
        assert rel_fro(got, want) <= TOL_FP32
        y0 = xmat(40000, dim, seed=1)
        out = dev(y0)
        capi.spmm(dev(x), dev(rp), dev(ci), out=out, accumulate=True, precision="fp32")
        assert rel_fro(out.cpu().numpy(), want + y0) <= TOL_FP32
        old = capi.set_tuning("pad_odd", 0)
        try:
            got = capi.spmm(dev(x), dev(rp), dev(ci), precision="fp32").cpu().numpy()
        finally:
            capi.set_tuning("pad_odd", old)
        assert rel_fro(got, want) <= TOL_FP32


# ---- tcgen05 dense super-window path -----------------------------------------------------------
@pytest.mark.parametrize("warp_specialised", [0, 1, 2, 3])
@pytest.mark.parametrize("name", ["sbm_1024", "rmat_1000", "band2_320", "holes_777", "dense_2048"])
def test_spmm_dense_superwindows_tcgen05(capi, name, warp_specialised):
    """warp_specialised: 2 = five-role kernel of dense_tma.cu with cp.async gathers (default), 3 = the same kernel
    with TMA gather4, 1 / 0 = the kernels of dense.cu."""
    rp, ci = GRAPHS[name]
    n = rp.size - 1
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "all_tc")
    old = capi.set_tuning("umma", 1)
    old_ws = capi.set_tuning("dense_ws", min(warp_specialised, 1))
    old_tma = capi.set_tuning("dense_tma", {2: 2, 3: 3}.get(warp_specialised, 0))
    try:
        plan = capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=0.0)
        assert plan.n_dense == sum(1 for s in range((n + 127) // 128) if rp[min(128 * s + 128, n)] > rp[128 * s])
        for dim in (16, 32, 64, 128, 256, 48):
            x = xmat(x_rows_for(rp, ci), dim, seed=dim + 3)
            fp32 = oracle.spmm(rp, ci, x, precision=1)
            tf32 = oracle.spmm(rp, ci, oracle.tf32_round(x), precision=1)
            got = capi.spmm_plan(dev(x), d_rp, d_ci, bp, etc, etr, ht, plan).cpu().numpy()
            assert capi.lib().hcspmm_debug_umma_error() == 0
            assert rel_fro(got, fp32) <= TOL_TF32, (name, dim)
            assert rel_fro(got, tf32) <= 2e-5, (name, dim)
            acc = dev(np.ones_like(fp32))
            capi.spmm_plan(dev(x), d_rp, d_ci, bp, etc, etr, ht, plan, out=acc, accumulate=True)
            assert rel_fro(acc.cpu().numpy(), tf32 + 1.0) <= 2e-5, (name, dim)
        # widths the dense kernel does not take fall back to the per-window paths
        x = xmat(x_rows_for(rp, ci), 40, seed=1)
        got = capi.spmm_plan(dev(x), d_rp, d_ci, bp, etc, etr, ht, plan).cpu().numpy()
        assert rel_fro(got, oracle.spmm(rp, ci, x, precision=1)) <= TOL_TF32
    finally:
        capi.set_tuning("umma", old)
        capi.set_tuning("dense_ws", old_ws)
        capi.set_tuning("dense_tma", old_tma)


@pytest.mark.parametrize("dim,hidden", [(64, 32), (128, 128), (256, 256), (48, 47), (16, 256), (256, 16), (96, 100)])
@pytest.mark.parametrize("name", ["sbm_1024", "dense_2048"])
@pytest.mark.parametrize("gather", [2, 3])
def test_fused_aggregation_update_on_tcgen05(capi, name, dim, hidden, gather):
    """hcspmm_spmm_gemm_aux with a plan that covers the graph: ONE kernel computes Z = A X (TMA gather4 + tcgen05) and
    out = rna(Z) rna(W) from the TMEM-resident aggregate.  Z equals the unfused dense path bit for bit; out equals
    the oracle's TF32 GEMM of that Z."""
    rp, ci = GRAPHS[name]
    n = rp.size - 1
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "all_tc")
    plan = capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=0.0)
    aux = capi.GraphAux(d_rp, d_ci, ht, plan)
    assert aux.plan_full == 1
    x, w = xmat(n, dim, seed=dim), xmat(dim, hidden, seed=hidden + 1)
    old_g = capi.set_tuning("dense_tma", gather)                # 2: cp.async gathers, 3: TMA gather4
    try:
        out, z = capi.spmm_gemm_aux(dev(x), d_rp, d_ci, bp, etc, etr, ht, dev(w), aux)
        assert capi.lib().hcspmm_debug_umma_error() == 0
        z_plain = capi.spmm_aux(dev(x), d_rp, d_ci, bp, etc, etr, ht, aux)
        assert torch.equal(z, z_plain)
        tf32 = oracle.spmm(rp, ci, oracle.tf32_round(x), precision=1)
        assert rel_fro(z.cpu().numpy(), tf32) <= 2e-5
        assert rel_fro(out.cpu().numpy(), oracle.gemm(z.cpu().numpy(), w, tf32=True)) <= 1e-4
        old = capi.set_tuning("fuse_update", 0)                     # the unfused route gives the same numbers
        try:
            out2, z2 = capi.spmm_gemm_aux(dev(x), d_rp, d_ci, bp, etc, etr, ht, dev(w), aux)
        finally:
            capi.set_tuning("fuse_update", old)
        assert torch.equal(z2, z) and rel_fro(out2.cpu().numpy(), out.cpu().numpy()) <= 1e-5
    finally:
        capi.set_tuning("dense_tma", old_g)


def test_fused_update_falls_back_when_the_plan_is_partial(capi):
    rp, ci = GRAPHS["sbm_1024"]
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, _ = capi.preprocess(d_ci, d_rp, "all_tc")
    ht = np.ones(64, np.int32)
    ht[40:48] = 0                                               # super-window 5 stays on the CUDA cores
    plan = capi.DensePlan(d_rp, d_ci, etr, dev(ht), min_reuse=0.0)
    aux = capi.GraphAux(d_rp, d_ci, dev(ht), plan)
    assert plan.n_dense == 7 and aux.plan_full == 0
    x, w = xmat(1024, 64, seed=2), xmat(64, 32, seed=3)
    out, z = capi.spmm_gemm_aux(dev(x), d_rp, d_ci, bp, etc, etr, dev(ht), dev(w), aux)
    want = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0)
    assert rel_fro(z.cpu().numpy(), want) <= 2e-5
    assert rel_fro(out.cpu().numpy(), oracle.gemm(z.cpu().numpy(), w, tf32=True)) <= 1e-4


def test_dense_tma_rectangular_operand_ids_outside_x_contribute_zero(capi):
    """Column ids >= x_rows (row shards keep global ids) are gathered as out-of-bounds rows: the TMA unit fills zeros."""
    rp, ci = GRAPHS["sbm_1024"]
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "all_tc")
    plan = capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=0.0)
    x = xmat(1024, 64, seed=9)
    x_short = x[:900]
    x_zero = x.copy()
    x_zero[900:] = 0
    got = capi.spmm_plan(dev(x_short), d_rp, d_ci, bp, etc, etr, ht, plan).cpu().numpy()
    assert capi.lib().hcspmm_debug_umma_error() == 0
    assert rel_fro(got, oracle.spmm(rp, ci, oracle.tf32_round(x_zero), precision=1)) <= 2e-5


def test_spmm_dense_plan_mixed_labels(capi):
    """Only super-windows whose eight windows are ALL tensor-core go dense; the rest stay hybrid."""
    rp, ci = GRAPHS["sbm_1024"]
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, _ = capi.preprocess(d_ci, d_rp, "all_tc")
    ht = np.ones(64, np.int32)
    ht[3] = 0            # super-window 0 has a CUDA-core window
    ht[40:48] = 0        # super-window 5 entirely CUDA-core
    old = capi.set_tuning("umma", 1)
    try:
        plan = capi.DensePlan(d_rp, d_ci, etr, dev(ht), min_reuse=0.0)
        assert plan.n_dense == 6
        x = xmat(1024, 64, seed=5)
        got = capi.spmm_plan(dev(x), d_rp, d_ci, bp, etc, etr, dev(ht), plan).cpu().numpy()
        want = oracle.spmm(rp, ci, x, hybrid_type=ht, precision=0)
        assert rel_fro(got, want) <= 2e-5
        assert capi.lib().hcspmm_debug_umma_error() == 0
        # b200-style reuse threshold: a huge threshold selects nothing
        assert capi.DensePlan(d_rp, d_ci, etr, dev(ht), min_reuse=1e6).n_dense == 0
    finally:
        capi.set_tuning("umma", old)


def test_b200_selector_labels_candidates_and_wide_dense(capi):
    """`b200` labels tensor-core CANDIDATES (3): without a dense plan they are exact FP32 CUDA-core windows;
    with one, the covered super-windows are TF32 on tcgen05 -- also for dim > 256 (column blocks of 256)."""
    rp, ci = GRAPHS["sbm_1024"]
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "b200")
    labels = ht.cpu().numpy()
    assert set(np.unique(labels)) <= {0, 3} and (labels == 3).any()
    old = capi.set_tuning("umma", 1)
    try:
        plan = capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=2.0)
        assert plan.n_dense > 0
        for dim in (64, 256, 384, 512):
            x = xmat(1024, dim, seed=dim)
            fp32 = oracle.spmm(rp, ci, x, precision=1)
            got = capi.spmm(dev(x), d_rp, d_ci, bp, etc, etr, ht).cpu().numpy()          # no plan: all CUDA-core
            assert rel_fro(got, fp32) <= TOL_FP32, dim
            got = capi.spmm_plan(dev(x), d_rp, d_ci, bp, etc, etr, ht, plan).cpu().numpy()
            assert rel_fro(got, fp32) <= TOL_TF32, dim
            assert rel_fro(got, fp32) > 1e-6, dim                                        # the tensor cores did run (TF32)
            assert capi.lib().hcspmm_debug_umma_error() == 0
    finally:
        capi.set_tuning("umma", old)


def test_b200_refit_rules(capi):
    """The re-fitted selector (benchmarks/selector_fit.py): (a) `b200_window` labels a 16-row window tensor-core iff the
    reference's logistic form with the B200 coefficients says so; (b) a super-window joins the dense plan only when its
    rows hold >= 8 entries on average (knob dense_min_rowlen)."""
    for name in ("sbm_1024", "rmat_hub_4096", "dense_2048", "band2_320"):
        rp, ci = GRAPHS[name]
        n = rp.size - 1
        bp, etc, etr, ht = (t.cpu().numpy() for t in capi.preprocess(dev(ci), dev(rp), "b200_window"))
        want = np.zeros_like(ht)
        for w in range((n + 15) // 16):
            e0, e1 = rp[16 * w], rp[min(16 * w + 16, n)]
            if e1 == e0:
                continue
            u = int(etc[e0:e1].max()) + 1
            dens = np.float32(e1 - e0) / np.float32(bp[w] * 128)
            z = float(u - 1) * -0.02312523 + float(dens) * -9.74306426 + 4.93743285
            want[w] = int(not (z > 0.0) and bp[w] * 8 <= 1024)
        assert np.array_equal(ht, want), name
    rp, ci = GRAPHS["band2_320"]                       # 4 entries per row, every column shared by 5 rows
    d_rp, d_ci = dev(rp), dev(ci)
    bp, etc, etr, ht = capi.preprocess(d_ci, d_rp, "all_tc")
    assert capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=2.0).n_dense == 0           # rows too short for tcgen05 to pay
    old = capi.set_tuning("dense_min_rowlen", 1)
    try:
        assert capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=2.0).n_dense == 3
    finally:
        capi.set_tuning("dense_min_rowlen", old)
    assert capi.DensePlan(d_rp, d_ci, etr, ht, min_reuse=0.0).n_dense == 3           # 0 = force


# ---- BF16-stored X --------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [8, 32, 64, 128, 256, 512, 200, 100])
@pytest.mark.parametrize("name", ["rmat_1000", "rmat_hub_4096", "ring3_256", "holes_777"])
def test_spmm_bf16_storage(capi, name, dim):
    rp, ci = GRAPHS[name]
    x = xmat(x_rows_for(rp, ci), dim, seed=dim + 11)
    fp32 = oracle.spmm(rp, ci, x, precision=1)
    bf16 = oracle.spmm(rp, ci, x, precision=2)          # X rounded to bfloat16 (RNE), FP32 sum
    got = capi.spmm(dev(x), dev(rp), dev(ci), precision="bf16").cpu().numpy()
    assert rel_fro(got, fp32) <= 1e-2                   # the north star's BF16 bar
    if dim % 8 == 0:
        assert rel_fro(got, bf16) <= 2e-5               # same rounding as the oracle's restatement
    else:
        assert rel_fro(got, fp32) <= TOL_FP32           # odd widths are computed in FP32
