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


SEG_SHIFT = 29          # bits 29..31 of a segment-tagged column id: the X buffer the row is read from
SEG_MASK = (1 << SEG_SHIFT) - 1


def pulled_layout(colidx64: torch.Tensor, bounds: torch.Tensor, rank: int, max_refs: int):
    """Segment mode of the peer exchange: which rows of X rank `rank` keeps in its operand.
    colidx64: the shard's global column ids; bounds: the row cuts (int64 tensor, world + 1 entries).
    -> (ids, cold): ids = the operand layout [pulled rows of lower ranks | ALL own rows | pulled rows of higher ranks],
    ascending global ids in each part, where a remote row is pulled iff the shard references it more than max_refs
    times; cold = the remote rows left at their owner (read in place by the gather), with their reference counts."""
    uniq, refs = torch.unique(colidx64, return_counts=True)
    own = torch.bucketize(uniq, bounds[1:-1], right=True)
    cold = (own != rank) & (refs <= max_refs)
    keep = ~cold
    mine = torch.arange(int(bounds[rank]), int(bounds[rank + 1]), device=colidx64.device, dtype=torch.int64)
    ids = torch.cat([uniq[keep & (own < rank)], mine, uniq[keep & (own > rank)]])
    return ids, (uniq[cold], refs[cold])


def tag_segments(colidx64: torch.Tensor, ids: torch.Tensor, bounds: torch.Tensor, rank: int, world: int,
                 firsts) -> torch.Tensor:
    """Segment-tagged int32 column ids of a shard (hcspmm_aux_t.d_colidx_segments): an entry whose row is in the local
    operand `ids` -> its position there (segment 0); any other entry -> segment (owner - rank) mod world in bits 29..31
    and, below, the row inside the OWNER's operand: firsts[owner] (where the owner's own rows start in its operand) +
    the row's index at its owner."""
    assert world <= 8 and ids.numel() <= SEG_MASK
    pos = torch.searchsorted(ids, colidx64).clamp_(max=max(0, ids.numel() - 1))
    local = ids[pos] == colidx64
    owner = torch.bucketize(colidx64, bounds[1:-1], right=True)
    first_t = torch.as_tensor(list(firsts), device=colidx64.device, dtype=torch.int64)
    row = first_t[owner] + colidx64 - bounds[owner]
    assert int(row.max()) <= SEG_MASK if row.numel() else True
    far = (((owner - rank) % world) << SEG_SHIFT) | row
    v = torch.where(local, pos, far)
    return torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()


def row_block_order(rowptr: torch.Tensor, colidx_local: torch.Tensor, n_operand_rows: int, n_blocks: int):
    """Row-block pipeline of the peer exchange: cut a shard's rows into n_blocks contiguous blocks of about equal
    stored entries (16-row aligned) and find, for every row of the exchange operand, the FIRST block that references
    it.  colidx_local: operand row of every entry, or a negative / out-of-range value for entries that do not read the
    local operand (rows read in place from a peer).  -> (cuts: n_blocks + 1 row offsets, first: int64[n_operand_rows],
    n_blocks where a row is never referenced).  Block b's SpMM can start once the operand rows with first <= b are in
    place."""
    n = rowptr.numel() - 1
    cuts = window_cuts(rowptr, n_blocks)
    dev = colidx_local.device
    rp = rowptr.to(torch.int64)
    ent_cuts = rp[torch.tensor(cuts, device=rp.device)]
    pos = torch.arange(colidx_local.numel(), device=dev, dtype=torch.int64)
    blk = torch.bucketize(pos, ent_cuts[1:-1].to(dev), right=True)               # block of every entry
    c = colidx_local.to(torch.int64)
    ok = (c >= 0) & (c < n_operand_rows)
    first = torch.full((n_operand_rows,), n_blocks, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, c[ok], blk[ok], reduce="amin", include_self=True)
    return cuts, first
