"""GPU tests of the NVLink peer-memory exchange pieces (csrc/peer.cu) that need only ONE device: the
halo-pull kernel against torch indexing (three local buffers stand in for three owners), the flag
barrier at world size 1, and IPC buffer allocation.  The multi-process path (CUDA IPC between ranks)
is exercised by scripts/gpu_peer_check.py under torchrun on a multi-GPU box."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,col0,width", [(128, 0, 128), (128, 32, 64), (48, 0, 48), (256, 128, 128), (4, 0, 4)])
def test_halo_pull_matches_indexing(dim, col0, width):
    from hcspmm import peer
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(dim + col0)
    owners = [torch.randn(n, dim, device=dev, generator=g) for n in (300, 17, 1000)]
    counts = [123, 17, 777]
    src = [torch.sort(torch.randperm(o.shape[0], device=dev, generator=g)[:c]).values for o, c in zip(owners, counts)]
    src_row = torch.cat(src).to(torch.int32)
    seg = torch.tensor([0, 123, 140, 917], dtype=torch.int32, device=dev)
    table = torch.tensor([o.data_ptr() for o in owners], dtype=torch.int64, device=dev)
    dst = torch.full((917, dim), -1.0, device=dev)
    # the caller's own segment (here: owner 1, rows [123, 140)) is skipped
    peer.halo_pull(table, dim, src_row, seg, 3, dst, col0, width, owner_mask=0b101, first_owner=2)
    want = torch.cat([o[i.long()] for o, i in zip(owners, src)])
    assert bool((dst[123:140] == -1.0).all())
    dst[123:140, col0:col0 + width] = want[123:140, col0:col0 + width]
    assert torch.equal(dst[:, col0:col0 + width], want[:, col0:col0 + width])
    dst2 = torch.full((917, dim), -1.0, device=dev)
    peer.halo_pull(table, dim, src_row, seg, 3, dst2, col0, width)          # every segment
    assert torch.equal(dst2[:, col0:col0 + width], want[:, col0:col0 + width])
    untouched = torch.ones(dim, dtype=torch.bool, device=dev)
    untouched[col0:col0 + width] = False
    assert bool((dst[:, untouched] == -1.0).all())


def test_halo_pull_subset_with_destination_rows():
    """hcspmm_halo_pull_rows: a subset of the halo (the part one row block needs first) lands in the operand rows the
    list names; rows outside the list stay untouched."""
    from hcspmm import peer
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    dim = 128
    owners = [torch.randn(n, dim, device=dev, generator=g) for n in (300, 40, 1000)]
    counts = [123, 0, 777]                                                     # owner 1 is the caller: nothing pulled
    src = [torch.sort(torch.randperm(o.shape[0], device=dev, generator=g)[:c]).values for o, c in zip(owners, counts)]
    seg_all = [0, 123, 163, 940]                                               # operand: [owner 0 | 40 own rows | owner 2]
    pos = torch.cat([torch.arange(0, 123, device=dev), torch.arange(163, 940, device=dev)])
    src_all = torch.cat([src[0], src[2]])
    pick = torch.sort(torch.randperm(900, device=dev, generator=g)[:300]).values     # this block's part of the halo
    dst_row, src_row = pos[pick].to(torch.int32), src_all[pick].to(torch.int32)
    seg = torch.searchsorted(dst_row.long(), torch.tensor(seg_all, device=dev)).to(torch.int32)
    table = torch.tensor([o.data_ptr() for o in owners], dtype=torch.int64, device=dev)
    dst = torch.full((940, dim), -1.0, device=dev)
    peer.halo_pull(table, dim, src_row, seg, 3, dst, 0, dim, owner_mask=0b101, first_owner=2, dst_row=dst_row)
    want = torch.full((940, dim), -1.0, device=dev)
    full = torch.cat([owners[0][src[0]], torch.full((40, dim), -1.0, device=dev), owners[2][src[2]]])
    want[dst_row.long()] = full[dst_row.long()]
    assert torch.equal(dst, want)


def test_peer_memory_single_rank_barrier_and_buffers():
    from hcspmm import peer
    dev = torch.device("cuda", 0)
    pm = peer.PeerMemory(dev)
    ptr, ptrs = pm.shared(1024 * 4)
    assert ptrs == [ptr]
    t = pm.tensor(ptr, (32, 32))
    assert float(t.abs().sum()) == 0.0           # zero-filled
    t.fill_(3.0)
    for _ in range(3):
        pm.barrier()
    torch.cuda.synchronize()
    pm.check()
    assert pm.epoch == 3 and float(pm.tensor(ptr, (1024,)).sum()) == 3072.0
    del t
    pm.close()


def test_halo_pull_rejects_misaligned():
    from hcspmm import capi
    L = capi.lib()
    d = torch.zeros(8, 6, device="cuda")
    rc = L.hcspmm_halo_pull(d.data_ptr(), 6, d.data_ptr(), d.data_ptr(), 1, 1, 0, 8, 0, 6, d.data_ptr(), 6, None)
    assert rc == -2


@pytest.mark.parametrize("dim,col0,width", [(128, 0, 128), (64, 16, 32), (256, 0, 256)])
def test_halo_push_matches_indexing(dim, col0, width):
    """The owner-side push: rows send_row[j] of the local shard land in each consumer's operand segment, in list
    order; three local buffers stand in for three peers, the caller's own bit is left clear."""
    from hcspmm import peer
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(dim + col0)
    src = torch.randn(500, dim, device=dev, generator=g)
    counts = [123, 0, 77]                                        # what peers 0 / 1 (self) / 2 want
    lists = [torch.sort(torch.randperm(500, device=dev, generator=g)[:c]).values for c in counts]
    send_row = torch.cat(lists).to(torch.int32)
    send_seg = torch.tensor([0, 123, 123, 200], dtype=torch.int32, device=dev)
    peers = [torch.full((300, dim), -1.0, device=dev) for _ in range(3)]
    offs = [40, 0, 11]                                           # where this rank's segment starts in each operand
    table = torch.tensor([p.data_ptr() + o * dim * 4 for p, o in zip(peers, offs)], dtype=torch.int64, device=dev)
    peer.halo_push(src, send_row, send_seg, table, dim, 3, 0b101, first_peer=2, col0=col0, width=width)
    torch.cuda.synchronize()
    for s in (0, 2):
        got = peers[s][offs[s]: offs[s] + counts[s], col0:col0 + width]
        assert torch.equal(got, src[lists[s].long(), col0:col0 + width])
        rest = peers[s].clone()
        rest[offs[s]: offs[s] + counts[s], col0:col0 + width] = -1.0
        assert bool((rest == -1.0).all())                        # nothing else was written
    assert bool((peers[1] == -1.0).all())


@pytest.mark.parametrize("dim,bf16", [(128, False), (32, False), (256, False), (64, True), (128, True), (512, True)])
@pytest.mark.parametrize("graph", ["rmat_hub_4096", "products_like"])
def test_spmm_segments_matches_one_buffer(graph, dim, bf16):
    """Segment mode of the balanced kernel (hcspmm_aux_t.d_colidx_segments) on ONE device: the rows of X are dealt to
    eight buffers (stand-ins for the local operand and seven peers' operands, different sizes, rows offset inside
    them), column ids are tagged with the buffer of their row -- the result equals the plain aggregation of the same
    X bit for bit (same loads, same sum order) in FP32 and in BF16-stored mode."""
    import numpy as np
    import HCSPMM
    from helpers import small_graphs
    from hcspmm import graphs
    dev = torch.device("cuda", 0)
    if graph == "products_like":
        rp, ci = graphs.rmat(40_000, 1_000_000, seed=11)
        rp, ci = rp.numpy(), ci.numpy()
    else:
        rp, ci = small_graphs()[graph]
    n = rp.size - 1
    d_rp, d_ci = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
    g = torch.Generator(device=dev).manual_seed(dim)
    x = torch.randn(n, dim, device=dev, generator=g)
    HCSPMM.set_classifier("shipped")
    old_sort = HCSPMM.set_row_sort(False)       # segment mode walks the CSR in place: compare with the same walk
    try:
        pre = HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16)
    finally:
        HCSPMM.set_row_sort(old_sort)
    xs = x.to(torch.bfloat16) if bf16 else x
    want = torch.empty(n, dim, device=dev)
    if bf16:
        HCSPMM.spmm_bf16(xs, d_rp, d_ci, *pre[:4], want, False, *pre[4:6])
    else:
        want = HCSPMM.forward(x, d_rp, d_ci, *pre)[0]
    # deal the rows: segment of row r, and its position inside that segment's buffer (after `lead` foreign rows)
    seg_of = torch.randint(0, 8, (n,), device=dev, generator=g)
    seg_of[: n // 2] = 0                                   # half of X stays local, like a real operand
    leads = [0, 5, 0, 17, 3, 0, 9, 1]
    bufs, row_in = [], torch.empty(n, dtype=torch.int64, device=dev)
    for s_ in range(8):
        idx = torch.nonzero(seg_of == s_).flatten()
        b = torch.full((leads[s_] + idx.numel() + 2, dim), float("nan"), device=dev).to(xs.dtype)
        b[leads[s_]: leads[s_] + idx.numel()] = xs[idx]
        row_in[idx] = leads[s_] + torch.arange(idx.numel(), device=dev)
        bufs.append(b)
    c64 = d_ci.long()
    v = (seg_of[c64] << 29) | row_in[c64]
    tagged = torch.where(v >= 2 ** 31, v - 2 ** 32, v).to(torch.int32).contiguous()
    x_rows = max(b.shape[0] for b in bufs)
    out = torch.empty(n, dim, device=dev)
    step = 256
    for c0 in range(0, dim, step):
        segs = [0] + [b.data_ptr() + c0 * b.element_size() for b in bufs[1:]]
        HCSPMM.spmm_segments(bufs[0][:, c0:c0 + step], d_rp, tagged, segs, x_rows, out[:, c0:c0 + step], False, pre[4], pre[5])
    torch.cuda.synchronize()
    assert not bool(torch.isnan(out).any())
    if dim <= 256:
        assert torch.equal(out, want)
    else:       # column blocks of 256 use another item size than one 512-wide launch: same sums, other piece order
        assert float((out - want).norm() / want.norm()) <= 1e-6


def test_spmm_segments_rejects_dense_plans_and_ignores_labels():
    """Segment mode runs on the CUDA-core balanced kernel in FP32: window labels do not matter (the result is the exact
    aggregation), a dense super-window plan cannot be combined with it and is refused loudly."""
    import HCSPMM
    from helpers import small_graphs
    dev = torch.device("cuda", 0)
    rp, ci = small_graphs()["sbm_1024"]
    n = rp.size - 1
    d_rp, d_ci = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
    x = torch.randn(n, 64, device=dev)
    HCSPMM.set_classifier("shipped")
    pre0 = HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16)
    HCSPMM.set_precision("fp32")
    try:
        want = HCSPMM.forward(x, d_rp, d_ci, *HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16))[0]
    finally:
        HCSPMM.set_precision("tf32")
    HCSPMM.set_classifier("all_tc")
    try:
        pre_tc = HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16)
        HCSPMM.set_dense(True)
        pre_dense = HCSPMM.preprocess(d_ci, d_rp, n, d_ci.numel(), (n + 15) // 16)
    finally:
        HCSPMM.set_classifier("shipped")
        HCSPMM.set_dense(False)
    assert int(pre_dense[4][1]) > 0
    got = HCSPMM.spmm_segments(x, d_rp, d_ci, [0] * 8, n, torch.empty(n, 64, device=dev), False, pre_tc[4], pre_tc[5])
    assert float((got - want).norm() / want.norm()) <= 1e-6
    with pytest.raises(RuntimeError, match="segment"):
        HCSPMM.spmm_segments(x, d_rp, d_ci, [0] * 8, n, torch.empty(n, 64, device=dev), False, pre_dense[4], pre_dense[5])
    del pre0
