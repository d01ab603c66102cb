"""GPU parity at BASELINE.json's FULL sizes (Reddit shape 115 M entries D = 256, products shape 62 M entries
D = 128, proteins shape with the tcgen05 dense plan) through size-independent properties -- the CPU oracle
cannot finish these shapes in seconds:
  * exact integer results:  A * 1 = row degrees;  A * [indicator columns] = |N(i) ∩ S_j|  (integers < 2^24 are
    exact in FP32 whatever the summation order), checked against torch index arithmetic on the CSR;
  * linearity:  A (a X1 + b X2) = a A X1 + b A X2;
  * checksum of checksums:  1^T (A X) = deg^T X  (the shapes are symmetric);
  * preprocessing:  edgeToRow = the row of every entry; blockPartition = ceil(#distinct columns / 8) per window
    (torch.unique on (window, column) pairs); 0 <= edgeToColumn < 8 * blockPartition; shipped labels all 0.
Everything goes through the C ABI (hcspmm.capi)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from hcspmm import capi as c
    c.lib()
    return c


def _rows_of(rp):
    n = rp.numel() - 1
    return torch.repeat_interleave(torch.arange(n, device=rp.device), (rp[1:] - rp[:-1]).long())


@pytest.mark.parametrize("shape,dim", [("reddit", 256), ("products", 128)])
def test_fullsize_spmm_properties(capi, shape, dim):
    from hcspmm import graphs
    dev = torch.device("cuda", 0)
    rp, ci, info = graphs.named(shape, device=dev)
    n, nnz = info["n"], info["nnz"]
    deg = (rp[1:] - rp[:-1]).float()
    rows = _rows_of(rp)

    # exact: degrees and neighbour counts in 7 vertex classes (+ one all-ones column)
    ind = torch.zeros(n, 8, device=dev)
    ind[:, 0] = 1.0
    for j in range(1, 8):
        ind[:, j] = ((torch.arange(n, device=dev) * 2654435761 >> 7) % 7 == j - 1).float()
    got = capi.spmm(ind, rp, ci, precision="fp32")
    want = torch.zeros(n, 8, device=dev).index_add_(0, rows, ind[ci.long()])
    assert torch.equal(got[:, 0], deg)
    assert torch.equal(got, want)
    assert float(got[:, 1:].double().sum()) == float(nnz)  # the 7 classes partition the columns
    del ind, want, got

    g = torch.Generator(device=dev).manual_seed(7)
    x1 = torch.randn(n, dim, device=dev, generator=g)
    x2 = torch.randn(n, dim, device=dev, generator=g)
    y1 = capi.spmm(x1, rp, ci, precision="fp32")
    y2 = capi.spmm(x2, rp, ci, precision="fp32")
    y12 = capi.spmm(0.5 * x1 - 2.0 * x2, rp, ci, precision="fp32")
    lin = 0.5 * y1 - 2.0 * y2
    assert float((y12 - lin).norm() / lin.norm()) <= 1e-5
    # checksum of checksums (A symmetric): column sums of Y against deg^T X, in float64
    cs = y1.double().sum(0)
    ref = (deg.double().unsqueeze(1) * x1.double()).sum(0)
    assert float((cs - ref).abs().max() / ref.abs().max()) <= 1e-5
    # the drop-in module's default path (TF32 labels, shipped selector: all CUDA-core) gives the same numbers
    import HCSPMM
    HCSPMM.set_classifier("shipped")
    pre = HCSPMM.preprocess(ci, rp, n, nnz, (n + 15) // 16)
    ym = HCSPMM.forward(x1, rp, ci, *pre)[0]
    assert float((ym - y1).norm() / y1.norm()) <= 1e-6
    # BF16-stored X: the north star's 1e-2 bar
    yb = capi.spmm(x1, rp, ci, precision="bf16")
    assert float((yb - y1).norm() / y1.norm()) <= 1e-2


@pytest.mark.parametrize("shape", ["reddit", "products"])
def test_fullsize_preprocess_properties(capi, shape):
    from hcspmm import graphs
    dev = torch.device("cuda", 0)
    rp, ci, info = graphs.named(shape, device=dev)
    n = info["n"]
    w = (n + 15) // 16
    bp, etc, etr, ht = capi.preprocess(ci, rp, "shipped")
    rows = _rows_of(rp)
    assert torch.equal(etr.long(), rows)
    assert not bool(ht.any())                              # hybrid_all_kernel.cu:262 labels every window CUDA-core
    win = rows // 16
    key = torch.unique(win * n + ci.long())                # distinct (window, column) pairs
    u = torch.bincount(key // n, minlength=w)
    assert torch.equal(bp.long(), (u + 7) // 8)
    assert int(etc.min()) >= 0 and bool((etc.long() < 8 * bp.long()[win]).all())
    # ranks are ascending in the column id inside a window: rank == position among the window's distinct columns
    first = torch.searchsorted(key, win * n)               # offset of each entry's window in the pair list
    pos = torch.searchsorted(key, win * n + ci.long())
    assert torch.equal(etc.long(), pos - first)


def test_fullsize_proteins_dense_plan(capi):
    """proteins shape, b200 selector + tcgen05 dense super-windows against the exact FP32 CUDA-core result."""
    from hcspmm import graphs
    dev = torch.device("cuda", 0)
    rp, ci, info = graphs.named("proteins", device=dev)
    n = info["n"]
    bp, etc, etr, ht = capi.preprocess(ci, rp, "b200")
    old = capi.set_tuning("umma", 1)
    try:
        plan = capi.DensePlan(rp, ci, etr, ht, min_reuse=2.0)
        assert plan.n_dense > 0.5 * ((n + 127) // 128)     # community structure: most super-windows are dense
        x = torch.randn(n, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        exact = capi.spmm(x, rp, ci, precision="fp32")
        got = capi.spmm_plan(x, rp, ci, bp, etc, etr, ht, plan)
        rel = float((got - exact).norm() / exact.norm())
        assert 1e-7 < rel <= 1e-3                          # TF32 on the tensor cores, inside the north star's bar
        assert capi.lib().hcspmm_debug_umma_error() == 0
        ones = torch.ones(n, 16, device=dev)
        assert torch.equal(capi.spmm_plan(ones, rp, ci, bp, etc, etr, ht, plan)[:, 0], (rp[1:] - rp[:-1]).float())
    finally:
        capi.set_tuning("umma", old)
