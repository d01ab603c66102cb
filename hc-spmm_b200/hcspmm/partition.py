"""Row-window partitioning of the adjacency for multi-GPU runs.

A is split into contiguous ranges of 16-row windows, one per rank, with the cut points chosen
on the prefix sum of per-window non-zeros (power-law graphs: balance nnz, not rows).  A shard
keeps GLOBAL column ids, so the local operator is rectangular (n_local x n_global) and its
preprocessing products are exactly the corresponding slices of the global ones -- window
contents do not change, labels stay bit-exact.  Pure host/torch index arithmetic, no kernels.
"""
from __future__ import annotations

import torch

BLK_H = 16


def window_cuts(rowptr: torch.Tensor, world: int) -> list[int]:
    """Row cut points r_0=0 <= r_1 <= ... <= r_world = n, each a multiple of 16 (except n),
    minimising the largest shard's nnz greedily on the window prefix sum."""
    n = rowptr.numel() - 1
    nw = (n + BLK_H - 1) // BLK_H
    rp = rowptr.to(torch.int64)
    wstart = rp[torch.arange(0, nw, device=rp.device) * BLK_H]          # nnz prefix at window starts
    total = int(rp[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        w = int(torch.searchsorted(wstart, torch.tensor(target, device=rp.device), right=False))
        # choose the closer of the two neighbouring window boundaries
        if w > 0 and w <= nw - 1 and abs(int(wstart[w - 1]) - target) <= abs(int(wstart[w]) - target):
            w -= 1
        w = max(w, cuts[-1] // BLK_H)
        cuts.append(min(w * BLK_H, n))
    cuts.append(n)
    return cuts


def local_shard(rowptr: torch.Tensor, colidx: torch.Tensor, r0: int, r1: int):
    """CSR of rows [r0, r1) with global column ids -> (rowptr_local int32[r1-r0+1], colidx_local)."""
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    rp = (rowptr[r0:r1 + 1].to(torch.int64) - e0).to(torch.int32).contiguous()
    return rp, colidx[e0:e1].contiguous()


def split_by_source(rowptr: torch.Tensor, colidx: torch.Tensor, cuts: list[int]):
    """Split a (local) CSR into one CSR per SOURCE shard: block s keeps the entries whose column id
    lies in [cuts[s], cuts[s+1]).  Used by the pipelined aggregation: the block for shard s can run
    as soon as X_s has arrived, accumulating into Y.  Rows keep their CSR (ascending) order."""
    n = rowptr.numel() - 1
    dev = colidx.device
    bounds = torch.tensor(cuts, device=dev, dtype=colidx.dtype)
    owner = torch.bucketize(colidx, bounds[1:-1], right=True)          # source shard of each entry
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (rowptr[1:] - rowptr[:-1]).to(torch.int64))
    out = []
    for s in range(len(cuts) - 1):
        m = owner == s
        cnt = torch.bincount(rows[m], minlength=n)
        rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(cnt, 0)
        out.append((rp.to(torch.int32), colidx[m].contiguous()))
    return out
