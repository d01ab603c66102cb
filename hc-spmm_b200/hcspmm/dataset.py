"""Graph input formats of the reference: mirror of /root/reference/dataset.py.

`HCSPMM_dataset(path, dim, num_class, load_from_txt=True, verbose=False)` keeps the reference's
constructor and attributes (dataset.py:8-40): num_nodes, num_edges, num_features, num_classes,
edge_index, avg_degree, avg_edgeSpan, column_index / row_pointers (int32 CSR), degrees, x, y and the
three masks.  The two wire formats are the reference's:
  * text: one "dst,src" pair per line, 1-based (dataset.py:51-53);
  * .npz with src_li, dst_li, num_nodes (dataset.py:73-77).
What differs is how they are read: the text file is parsed in one vectorised pass (the reference loops
over lines in Python, dataset.py:50-57) and the CSR is built on the device with a sort + unique
(the reference goes through scipy coo -> csr on the host, dataset.py:93-98).  The result is the
same canonical CSR: rows ascending, column ids ascending inside a row, duplicates merged, values
dropped (A is binary for every kernel).  Unlike the synthetic generators, nothing is symmetrised and
self loops are kept: the file is taken as it is, exactly like the reference.
"""
from __future__ import annotations

import time

import numpy as np
import torch


def read_txt(path: str):
    """-> (src int64[E], dst int64[E]) 0-based, in file order."""
    with open(path, "rb") as f:
        buf = f.read()
    flat = np.array(buf.replace(b",", b" ").split(), dtype=np.int64)
    if flat.size % 2:
        raise ValueError(f"{path}: odd number of integers -- expected 'dst,src' per line")
    pairs = flat.reshape(-1, 2)
    return pairs[:, 1] - 1, pairs[:, 0] - 1      # line is "dst,src" (dataset.py:51)


def read_npz(path: str):
    if not path.endswith(".npz"):
        raise ValueError("graph file must be a .npz file")      # dataset.py:70-71
    g = np.load(path)
    return np.asarray(g["src_li"], dtype=np.int64), np.asarray(g["dst_li"], dtype=np.int64), int(g["num_nodes"])


def write_npz(path: str, src, dst, num_nodes: int) -> None:
    np.savez(path, src_li=np.asarray(src), dst_li=np.asarray(dst), num_nodes=num_nodes)


def csr_from_edges(src: torch.Tensor, dst: torch.Tensor, num_nodes: int):
    """Canonical int32 CSR of the (src -> row, dst -> column) list: what scipy's
    coo_matrix((1, (src, dst))).tocsr() yields (dataset.py:93-98) with the values dropped."""
    key = torch.unique(src.to(torch.int64) * num_nodes + dst.to(torch.int64))       # sorted, duplicates merged
    rows = torch.div(key, num_nodes, rounding_mode="floor")
    cols = (key - rows * num_nodes).to(torch.int32)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=key.device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=num_nodes), 0)
    if int(rowptr[-1]) >= 2 ** 31:
        raise ValueError("more than 2^31-1 stored entries: the CSR is int32 like the reference's")
    return rowptr.to(torch.int32), cols


class HCSPMM_dataset(torch.nn.Module):
    """data loading for more graphs (dataset.py:8)"""

    def __init__(self, path, dim, num_class, load_from_txt=True, verbose=False, device=None):
        super().__init__()
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.load_from_txt = load_from_txt
        self.num_nodes = 0
        self.num_features = dim
        self.num_classes = num_class
        self.edge_index = None
        self.reorder_flag = False
        self.verbose_flag = verbose
        self.avg_degree = -1
        self.avg_edgeSpan = -1
        self.init_edges(path)
        self.init_embedding(dim)
        self.init_labels(num_class)
        n = self.num_nodes
        idx = torch.arange(n, device=self.device)
        self.train_mask = idx < int(n * 1)          # dataset.py:31-40: leading fractions 1 / 0.3 / 0.1
        self.val_mask = idx < int(n * 0.3)
        self.test_mask = idx < int(n * 0.1)

    def init_edges(self, path):
        start = time.perf_counter()
        if self.load_from_txt:
            src, dst = read_txt(path)
            self.num_nodes = int(max(src.max(initial=-1), dst.max(initial=-1))) + 1     # max id + 1 (dataset.py:60)
        else:
            src, dst, self.num_nodes = read_npz(path)
        self.num_edges = int(src.size)
        self.edge_index = np.stack([src, dst])
        if self.verbose_flag:
            print("# Loading ({}) {:.3f}s ".format("txt" if self.load_from_txt else "npz", time.perf_counter() - start))
        self.avg_degree = self.num_edges / max(1, self.num_nodes)
        self.avg_edgeSpan = float(np.mean(np.abs(src - dst))) if src.size else 0.0
        if self.verbose_flag:
            print("# nodes: {}".format(self.num_nodes))
            print("# avg_degree: {:.2f}".format(self.avg_degree))
            print("# avg_edgeSpan: {}".format(int(self.avg_edgeSpan)))
        start = time.perf_counter()
        self.row_pointers, self.column_index = csr_from_edges(torch.from_numpy(src).to(self.device),
                                                              torch.from_numpy(dst).to(self.device), self.num_nodes)
        if self.verbose_flag:
            print("# Build CSR (s): {:.3f}".format(time.perf_counter() - start))
        deg = (self.row_pointers[1:] - self.row_pointers[:-1]).to(torch.float32)
        self.degrees = torch.sqrt(torch.clamp(deg, min=1.0))      # config.py func: 0 -> 1 (dataset.py:104-105)

    def init_embedding(self, dim):
        self.x = torch.randn(self.num_nodes, dim, device=self.device)

    def init_labels(self, num_class):
        self.y = torch.ones(self.num_nodes, dtype=torch.long, device=self.device)
