"""CPU tests of the graph input formats: hcspmm.dataset against the route the reference
takes (dataset.py:43-105: Python line loop + scipy coo -> csr), on text and .npz files with duplicates,
self loops, isolated trailing vertices and unsorted lines."""
import numpy as np
import scipy.sparse as sp
import torch

from hcspmm import dataset, graphs


def reference_route(src, dst, n):
    """dataset.py:93-105 verbatim in spirit: coo -> csr (duplicates summed, indices sorted), values dropped."""
    csr = sp.coo_matrix((np.ones(len(src)), (src, dst)), shape=(n, n)).tocsr()
    csr.sum_duplicates()
    csr.sort_indices()
    deg = np.diff(csr.indptr)
    return csr.indptr.astype(np.int32), csr.indices.astype(np.int32), np.sqrt(np.where(deg > 0, deg, 1)).astype(np.float32)


def messy_edges(seed=0, n=300, e=4000):
    r = np.random.default_rng(seed)
    src, dst = r.integers(0, n - 5, e), r.integers(0, n - 5, e)      # last 5 vertices isolated
    src[:50], dst[:50] = src[50:100], dst[50:100]                    # duplicates
    src[100:120] = dst[100:120]                                      # self loops
    return src, dst, n


def test_txt_format_matches_reference_route(tmp_path):
    src, dst, n = messy_edges()
    path = tmp_path / "g.txt"
    with open(path, "w") as f:
        for s, d in zip(src, dst):
            f.write(f"{d + 1},{s + 1}\n")                             # "dst,src", 1-based
    ds = dataset.HCSPMM_dataset(str(path), 16, 7, load_from_txt=True, device="cpu")
    n_txt = int(max(src.max(), dst.max())) + 1                        # the text route sizes by the largest id
    rp, ci, deg = reference_route(src, dst, n_txt)
    assert ds.num_nodes == n_txt and ds.num_edges == len(src)
    assert np.array_equal(ds.row_pointers.numpy(), rp) and np.array_equal(ds.column_index.numpy(), ci)
    assert ds.row_pointers.dtype == torch.int32 and ds.column_index.dtype == torch.int32
    assert np.allclose(ds.degrees.numpy(), deg)
    assert np.array_equal(ds.edge_index, np.stack([src, dst]))
    assert ds.x.shape == (n_txt, 16) and ds.y.shape == (n_txt,) and int(ds.y.min()) == 1
    assert abs(ds.avg_degree - len(src) / n_txt) < 1e-12
    assert abs(ds.avg_edgeSpan - np.mean(np.abs(src - dst))) < 1e-9
    assert int(ds.train_mask.sum()) == n_txt and int(ds.val_mask.sum()) == int(n_txt * 0.3)
    assert int(ds.test_mask.sum()) == int(n_txt * 0.1) and bool(ds.val_mask[0]) and not bool(ds.val_mask[-1])


def test_npz_format_matches_reference_route(tmp_path):
    src, dst, n = messy_edges(seed=1)
    path = str(tmp_path / "g.npz")
    dataset.write_npz(path, src, dst, n)
    ds = dataset.HCSPMM_dataset(path, 8, 3, load_from_txt=False, device="cpu")
    rp, ci, _ = reference_route(src, dst, n)
    assert ds.num_nodes == n                                           # npz carries num_nodes: isolated tail kept
    assert np.array_equal(ds.row_pointers.numpy(), rp) and np.array_equal(ds.column_index.numpy(), ci)
    try:
        dataset.HCSPMM_dataset(str(tmp_path / "g.bin"), 8, 3, load_from_txt=False, device="cpu")
        assert False
    except ValueError:
        pass


def test_write_txt_round_trip(tmp_path):
    rp, ci = graphs.rmat(500, 6000, seed=4)
    path = str(tmp_path / "rt.txt")
    graphs.write_txt(path, rp, ci)
    ds = dataset.HCSPMM_dataset(path, 4, 2, device="cpu")
    n = ds.num_nodes
    assert torch.equal(ds.row_pointers, rp[: n + 1]) and torch.equal(ds.column_index, ci)
    assert int(rp[-1]) == int(rp[n])                                   # only isolated vertices may be cut off
