"""Generates tests/golden/*.npz -- committed known-answer vectors for the hot path.

The reference ships no tests or golden vectors, so these are produced here,
in the development container, from sources that are independent of our CUDA code:
  * preprocessing arrays: the C restatement (oracle/hcspmm_oracle.c); on the GPU box the same
    arrays are also compared with the recompiled UNMODIFIED reference extension
    (tests/test_gpu_module.py::test_reference_preprocess_matches);
  * y_fp32: torch.sparse.mm, FP32, CPU -- the north star's stated oracle;
  * loa_perm: the UNMODIFIED reference LOI.cpp (oracle/_ref/libloi_ref.so, built from
    /root/reference/LOI.cpp by oracle/Makefile) -- reorder_plus_new_direct + main's file order.
Run:  python oracle/make_golden.py     (needs /root/reference for the LOA vectors)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hc-spmm_b200"), os.path.join(ROOT, "tests")]

import oracle  # noqa: E402
from helpers import small_graphs, torch_sparse_ref  # noqa: E402

NAMES = ["ring3_256", "band2_320", "rmat_1000", "holes_777", "uniform_777", "rect_72x100000", "empty_48"]


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    graphs = small_graphs()
    for name in NAMES:
        rp, ci = graphs[name]
        n = rp.size - 1
        xr = max(n, int(ci.max()) + 1 if ci.size else n)
        x = np.random.default_rng(12345).standard_normal((xr, 32)).astype(np.float32)
        d = dict(rowptr=rp, colidx=ci, x_seed=np.int64(12345), x_shape=np.array(x.shape))
        for mode, tag in ((0, "shipped"), (1, "intended")):
            bp, etc, etr, ht = oracle.preprocess(ci, rp, mode)
            d.update({f"bp_{tag}": bp, f"ht_{tag}": ht})
            d.update(etc=etc, etr=etr)
        d["y_fp32"] = torch_sparse_ref(rp, ci, x) if ci.size else np.zeros((n, 32), np.float32)
        if n <= 2000 and xr == n and oracle.have_loi_ref():   # LOA needs a square graph
            perm, sizes, nfull = oracle.loa_reference(rp, ci)
            d.update(loa_perm=perm, loa_block_sizes=sizes, loa_full=np.int64(nfull))
        np.savez_compressed(os.path.join(out, f"{name}.npz"), **d)
        print(name, {k: getattr(v, "shape", v) for k, v in d.items()})


if __name__ == "__main__":
    main()
