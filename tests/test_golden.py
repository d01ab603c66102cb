"""Known-answer tests against the committed fixtures in tests/golden (made by oracle/make_golden.py):
CPU: the oracle reproduces them; GPU: the CUDA path, through the C ABI, reproduces them."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import rel_fro

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def load(path):
    d = np.load(path)
    x = np.random.default_rng(int(d["x_seed"])).standard_normal(tuple(d["x_shape"])).astype(np.float32)
    return d, x


def test_fixtures_present():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_reproduces_golden(path):
    d, x = load(path)
    for mode, tag in ((0, "shipped"), (1, "intended")):
        bp, etc, etr, ht = oracle.preprocess(d["colidx"], d["rowptr"], mode)
        assert np.array_equal(bp, d[f"bp_{tag}"]) and np.array_equal(ht, d[f"ht_{tag}"])
        assert np.array_equal(etc, d["etc"]) and np.array_equal(etr, d["etr"])
    assert rel_fro(oracle.spmm(d["rowptr"], d["colidx"], x, precision=1), d["y_fp32"]) <= 1e-6
    if "loa_perm" in d.files:      # produced by the UNMODIFIED reference LOI.cpp
        perm, sizes, nfull = oracle.loa(d["rowptr"], d["colidx"])
        assert np.array_equal(perm, d["loa_perm"]) and np.array_equal(sizes, d["loa_block_sizes"])
        assert nfull == int(d["loa_full"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_cuda_path_reproduces_golden(path):
    from hcspmm import capi
    d, x = load(path)
    rp, ci = torch.from_numpy(d["rowptr"]).cuda(), torch.from_numpy(d["colidx"]).cuda()
    for tag in ("shipped", "intended"):
        bp, etc, etr, ht = capi.preprocess(ci, rp, tag)
        assert np.array_equal(bp.cpu().numpy(), d[f"bp_{tag}"]) and np.array_equal(ht.cpu().numpy(), d[f"ht_{tag}"])
        assert np.array_equal(etc.cpu().numpy(), d["etc"]) and np.array_equal(etr.cpu().numpy(), d["etr"])
    xd = torch.from_numpy(x).cuda()
    y = capi.spmm(xd, rp, ci, precision="fp32").cpu().numpy()
    assert rel_fro(y, d["y_fp32"]) <= 1e-5                       # FP32 CUDA-core path
    pre = capi.preprocess(ci, rp, "all_tc")
    y = capi.spmm(xd, rp, ci, *pre, precision="tf32").cpu().numpy()
    assert rel_fro(y, d["y_fp32"]) <= 1e-3                       # TF32 tensor-core path (north star)
