"""Seeded synthetic graphs in the reference's CSR convention.

Every generator returns ``(row_pointers int32[N+1], column_index int32[nnz])`` as torch
tensors on ``device`` -- the two arrays ``dataset.py:93-103`` of the reference builds
(scipy ``coo -> csr``: sorted, de-duplicated rows, values dropped).  Graphs are
undirected (symmetrised), without self loops.  All generators are written with torch
ops only so the big shapes (10^8 stored entries) are built on the GPU in well under a
second; CPU and CUDA generators give different (each reproducible) streams.
"""
from __future__ import annotations

import torch


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def csr_from_pairs(u: torch.Tensor, v: torch.Tensor, n: int, symmetrize: bool = True):
    """Canonical CSR (int32) of the pair list, like scipy coo->csr + sum_duplicates with the
    values dropped (dataset.py:93-103).  Self loops are removed."""
    u = u.to(torch.int64)
    v = v.to(torch.int64)
    keep = u != v
    u, v = u[keep], v[keep]
    if symmetrize:
        u, v = torch.cat([u, v]), torch.cat([v, u])
    key = torch.unique(u * n + v)  # sorted
    del u, v
    rows = torch.div(key, n, rounding_mode="floor")
    cols = (key - rows * n).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=cols.device)
    rowptr[1:] = torch.cumsum(counts, 0)
    assert int(rowptr[-1]) < 2 ** 31, "nnz must fit int32 (reference uses int CSR)"
    return rowptr.to(torch.int32), cols


def ring_matching(n: int, seed: int = 0, device="cpu"):
    """C1 'example': ring + one random perfect matching => every degree is 3 (48 edges per
    16-row window, inside the reference kernels' 62-edge capacity, hybrid_all_kernel.cu:26)."""
    assert n % 2 == 0
    g = _gen(seed, device)
    i = torch.arange(n, device=device)
    while True:
        p = torch.randperm(n, generator=g, device=device)
        a, b = p[: n // 2], p[n // 2:]
        d = (a - b).abs()
        if not bool(((d == 1) | (d == n - 1)).any()):  # matching edge must not duplicate a ring edge
            break
    return csr_from_pairs(torch.cat([i, a]), torch.cat([(i + 1) % n, b]), n)


def banded(n: int, half_bandwidth: int = 2, device="cpu"):
    """C1 'example_band': neighbours i±1..i±k (no wrap).  k=2 gives <= 20 distinct columns per
    16-row window (<= 24 = MAX_BLK*8, hybrid_all_kernel.cu:23)."""
    i = torch.arange(n, device=device)
    us, vs = [], []
    for k in range(1, half_bandwidth + 1):
        us.append(i[:-k])
        vs.append(i[k:])
    return csr_from_pairs(torch.cat(us), torch.cat(vs), n)


def _rmat_pairs(scale: int, m: int, g, device, a: float, b: float, c: float):
    u = torch.zeros(m, dtype=torch.int64, device=device)
    v = torch.zeros(m, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(m, generator=g, device=device)
        ubit = r >= (a + b)
        vbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
        u = (u << 1) | ubit.to(torch.int64)
        v = (v << 1) | vbit.to(torch.int64)
    return u, v


def rmat(n: int, nnz_target: int, seed: int = 1, a: float = 0.57, b: float = 0.19,
         c: float = 0.19, permute: bool = True, device="cpu", exact: bool = True):
    """Power-law graph: R-MAT(a,b,c,1-a-b-c) at scale ceil(log2 n), ids folded mod n, then a
    seeded random relabelling.  ``nnz_target`` counts STORED CSR entries after symmetrisation
    (how Reddit's 114.6 M is counted); with ``exact`` the undirected pair set is topped up and
    trimmed to nnz_target // 2 pairs."""
    g = _gen(seed, device)
    scale = max(1, (n - 1).bit_length())
    want = nnz_target // 2
    keys = torch.empty(0, dtype=torch.int64, device=device)
    batch = int(want * 1.15) + 1024
    for _ in range(64):
        u, v = _rmat_pairs(scale, batch, g, device, a, b, c)
        u, v = u % n, v % n
        lo, hi = torch.minimum(u, v), torch.maximum(u, v)
        k = lo * n + hi
        k = k[lo != hi]
        keys = torch.unique(torch.cat([keys, k]))
        if keys.numel() >= want or not exact:
            break
        batch = max(1024, int((want - keys.numel()) * 1.5))
    if exact and keys.numel() > want:
        sel = torch.randperm(keys.numel(), generator=g, device=device)[:want]
        keys = keys[sel]
    u = torch.div(keys, n, rounding_mode="floor")
    v = keys - u * n
    del keys
    if permute:
        p = torch.randperm(n, generator=g, device=device)
        u, v = p[u], p[v]
    return csr_from_pairs(u, v, n)


def sbm_dense_windows(n: int, community: int = 512, p_in: float = 0.55, extra: int = 15,
                      seed: int = 3, device="cpu"):
    """C4 proteins-shape 'dense windows': communities of ``community`` consecutive vertices with
    intra-community edge probability p_in, plus ~``extra`` uniformly random neighbours per vertex,
    NO id permutation -- so a 16-row window condenses to community + a few hundred distinct
    columns at ~40 % tile density (the regime where the dense contraction wins)."""
    g = _gen(seed, device)
    us, vs = [], []
    ncomm = (n + community - 1) // community
    # intra-community: sample each unordered pair (i<j) with probability p_in, community by
    # community in batches to bound memory
    tri = torch.triu_indices(community, community, offset=1, device=device)
    per = tri.shape[1]
    batch = max(1, (1 << 26) // per)
    for c0 in range(0, ncomm, batch):
        c1 = min(ncomm, c0 + batch)
        r = torch.rand((c1 - c0, per), generator=g, device=device) < p_in
        ci, pi = r.nonzero(as_tuple=True)
        base = (ci + c0) * community
        uu, vv = base + tri[0][pi], base + tri[1][pi]
        ok = (uu < n) & (vv < n)
        us.append(uu[ok])
        vs.append(vv[ok])
    m = n * extra // 2
    us.append(torch.randint(0, n, (m,), generator=g, device=device))
    vs.append(torch.randint(0, n, (m,), generator=g, device=device))
    return csr_from_pairs(torch.cat(us), torch.cat(vs), n)


def banded_random(n: int, avg_degree: int, half_bandwidth: int, seed: int = 5, device="cpu"):
    """C5 banded graphs: each row's neighbours uniform in [i-b, i+b]."""
    g = _gen(seed, device)
    m = n * avg_degree // 2
    u = torch.randint(0, n, (m,), generator=g, device=device)
    off = torch.randint(-half_bandwidth, half_bandwidth + 1, (m,), generator=g, device=device)
    v = (u + off).clamp_(0, n - 1)
    return csr_from_pairs(u, v, n)


def uniform_random(n: int, avg_degree: int, seed: int = 7, device="cpu"):
    g = _gen(seed, device)
    m = n * avg_degree // 2
    u = torch.randint(0, n, (m,), generator=g, device=device)
    v = torch.randint(0, n, (m,), generator=g, device=device)
    return csr_from_pairs(u, v, n)


# Named shapes of BASELINE.json (N, stored nnz, feature width)
SHAPES = {
    "reddit": dict(n=232_965, nnz=114_615_892, dim=256, seed=1, kind="rmat"),
    "products": dict(n=2_449_029, nnz=61_859_140, dim=128, seed=2, kind="rmat"),
    "proteins": dict(n=132_534, nnz=39_561_252, dim=256, seed=3, kind="sbm"),
    # inside the reference kernels' validity envelope (<= 62 edges / window, dim 32, N % 16 == 0):
    # the one shape on which the recompiled reference extension can be timed beside ours
    "envelope": dict(n=1_048_576, nnz=3_145_728, dim=32, seed=0, kind="ring"),
}


def named(shape: str, device="cpu", scale: float = 1.0):
    """Build one of the BASELINE.json shapes (optionally scaled down by ``scale`` in both N and
    nnz, keeping the average degree) -> (rowptr, colidx, info)."""
    s = SHAPES[shape]
    n = max(64, int(s["n"] * scale) // 16 * 16 if scale != 1.0 else s["n"])
    nnz = int(s["nnz"] * scale)
    if s["kind"] == "rmat":
        rp, ci = rmat(n, nnz, seed=s["seed"], device=device)
    elif s["kind"] == "ring":
        rp, ci = ring_matching(n, seed=s["seed"], device=device)
    else:
        rp, ci = sbm_dense_windows(n, seed=s["seed"], device=device)
    return rp, ci, dict(shape=shape, n=n, nnz=int(rp[-1]), dim=s["dim"], scale=scale)


def write_txt(path: str, rowptr: torch.Tensor, colidx: torch.Tensor) -> None:
    """The reference's text format: one 'dst,src' pair per line, 1-based, sorted by src then dst
    (dataset.py:51-53 reads 'dst, src = line.split(",")'; LOI.cpp:493-499 needs src ascending)."""
    rp = rowptr.cpu().tolist()
    ci = colidx.cpu().tolist()
    with open(path, "w") as f:
        for r in range(len(rp) - 1):
            for e in range(rp[r], rp[r + 1]):
                f.write(f"{ci[e] + 1},{r + 1}\n")
