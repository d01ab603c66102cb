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
