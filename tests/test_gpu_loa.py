"""GPU tests of the LOA reordering kernel (A9): bit-exact with the CPU oracle (which is itself pinned
against the unmodified reference LOI.cpp, tests/test_oracle.py) and with the reference-produced
golden permutations in tests/golden."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import rel_fro, small_graphs
from hcspmm import graphs as G

pytestmark = pytest.mark.gpu
GRAPHS = small_graphs()


def gpu_loa(rp, ci):
    from hcspmm import capi
    perm, sizes, nfull = capi.loa_reorder(torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda())
    return perm.cpu().numpy(), sizes.numpy(), nfull


@pytest.mark.parametrize("name", ["ring3_256", "band2_320", "rmat_1000", "sbm_1024", "uniform_777", "holes_777",
                                  "empty_48", "single_row", "rmat_hub_4096"])
def test_loa_matches_oracle(name):
    rp, ci = GRAPHS[name]
    if name == "single_row":
        pytest.skip("not square")
    perm, sizes, nfull = gpu_loa(rp, ci)
    operm, osizes, ofull = oracle.loa(rp, ci)
    assert np.array_equal(sizes, osizes)
    assert nfull == ofull
    assert np.array_equal(perm, operm)


def test_loa_matches_reference_golden():
    """loa_perm in the fixtures was produced by the UNMODIFIED reference LOI.cpp."""
    paths = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
    seen = 0
    for path in paths:
        d = np.load(path)
        if "loa_perm" not in d.files:
            continue
        perm, sizes, nfull = gpu_loa(d["rowptr"], d["colidx"])
        assert np.array_equal(perm, d["loa_perm"]), path
        assert np.array_equal(sizes, d["loa_block_sizes"]) and nfull == int(d["loa_full"]), path
        seen += 1
    assert seen >= 4


def test_loa_larger_power_law():
    rp, ci = G.rmat(20000, 400000, seed=4)
    rp, ci = rp.numpy(), ci.numpy()
    perm, sizes, nfull = gpu_loa(rp, ci)
    operm, osizes, ofull = oracle.loa(rp, ci)
    assert np.array_equal(perm, operm) and nfull == ofull


def test_relabel_preserves_the_operator():
    """Y' = P A P^T (P X): relabelling with the LOA permutation permutes the SpMM result."""
    from hcspmm import capi
    rp, ci = GRAPHS["rmat_1000"]
    d_rp, d_ci = torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda()
    perm, _, _ = capi.loa_reorder(d_rp, d_ci)
    rp2, ci2 = capi.relabel(d_rp, d_ci, perm)
    assert rp2.numel() == d_rp.numel() and ci2.numel() == d_ci.numel()
    x = torch.randn(1000, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    y = capi.spmm(x, d_rp, d_ci, precision="fp32")
    y2 = capi.spmm(x[perm.long()].contiguous(), rp2, ci2, precision="fp32")
    assert rel_fro(y2.cpu().numpy(), y[perm.long()].cpu().numpy()) <= 1e-5
    # LOA's purpose: fewer distinct columns per 16-row window (more condensable tiles)
    bp_before = capi.preprocess(d_ci, d_rp)[0].sum().item()
    bp_after = capi.preprocess(ci2, rp2)[0].sum().item()
    assert bp_after <= bp_before
