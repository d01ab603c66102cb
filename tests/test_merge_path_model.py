"""CPU model of the balanced kernel's work decomposition (csrc/spmm.cu: merge_path_rows, spmm_balanced_kernel,
spmm_balanced_fixup_kernel): the same split search and the same head / tail / skip rules, restated in Python
and checked on random row-length patterns -- every row is written exactly once and equals its sum, whatever
the item size, including runs of empty rows, rows cut by many item boundaries and trailing empty rows.  The
CUDA implementation of these rules is checked on the GPU by tests/test_gpu_parity.py::test_spmm_balanced_items."""
import random

import numpy as np


def search(rowptr, n, nnz, diag):                       # merge_path_rows
    lo, hi = max(0, diag - nnz), min(diag, n)
    while lo < hi:
        mid = (lo + hi) // 2
        if rowptr[mid + 1] <= diag - 1 - mid:
            lo = mid + 1
        else:
            hi = mid
    return lo


def run(rowptr, vals, chunk):
    n, nnz = len(rowptr) - 1, int(rowptr[-1])
    total = n + nnz
    items = max(1, -(-total // chunk))
    y = np.full(n, np.nan)
    writes = np.zeros(n, int)
    part = np.full((items, 2), np.nan)
    split = np.full((items, 2), -1)
    for k in range(items):                              # spmm_balanced_kernel, one item
        d0, d1 = min(total, k * chunk), min(total, (k + 1) * chunk)
        x0, x1 = search(rowptr, n, nnz, d0), search(rowptr, n, nnz, d1)
        y0, y1 = d0 - x0, d1 - x1
        last_in = x1 < n and rowptr[x1] < y1
        rows = x1 - x0 + (1 if last_in else 0)
        if rows == 0:
            continue
        assert rows <= chunk + 1
        rp = [min(max(int(rowptr[x0 + i]), y0), y1) for i in range(rows + 1)]
        s0, t0, last = rowptr[x0], rowptr[x0 + 1], rows - 1
        is_tail = last_in and rowptr[x1 + 1] > y1
        head_skip = s0 < y0 and t0 <= y0
        is_head = s0 < y0 and t0 > y0 and not (is_tail and last == 0)
        split[k] = (x0 if is_head else -1, x1 if is_tail else -1)
        for i in range(rows):
            if i == 0 and head_skip:
                assert rp[1] == rp[0]
                continue
            s = vals[rp[i]:rp[i + 1]].sum()
            if i == 0 and is_head:
                assert rp[1] > rp[0]
                part[k, 0] = s
            elif i == last and is_tail:
                assert rp[i + 1] > rp[i]
                part[k, 1] = s
            else:
                assert rp[i] == rowptr[x0 + i] and rp[i + 1] == rowptr[x0 + i + 1]
                y[x0 + i] = s
                writes[x0 + i] += 1
    for k in range(items):                              # spmm_balanced_fixup_kernel
        r = split[k, 0]
        if r < 0:
            continue
        k1 = k
        while k1 > 0 and split[k1 - 1, 1] == r:
            k1 -= 1
        y[r] = sum(part[kk, 1] for kk in range(k1, k)) + part[k, 0]
        writes[r] += 1
    return y, writes


def test_every_row_written_once_and_correct():
    random.seed(1)
    rng = np.random.default_rng(1)
    for _ in range(1500):
        n = random.randint(1, 60)
        kind = random.random()
        if kind < 0.3:
            deg = [random.choice([0, 0, 0, 1, 2]) for _ in range(n)]
        elif kind < 0.6:
            deg = [random.choice([0, 1, 3, 40, 100]) for _ in range(n)]
        else:
            deg = [random.randint(0, 12) for _ in range(n)]
        rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(int)
        vals = rng.integers(1, 100, size=rowptr[-1]).astype(float)
        chunk = random.choice([1, 2, 3, 4, 7, 16, 64])
        y, writes = run(rowptr, vals, chunk)
        want = np.array([vals[rowptr[i]:rowptr[i + 1]].sum() for i in range(n)])
        assert (writes == 1).all(), (rowptr.tolist(), chunk)
        assert np.array_equal(y, want), (rowptr.tolist(), chunk)
