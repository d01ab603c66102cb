"""CPU tests of the multi-GPU host logic: nnz-balanced window cuts, shard extraction, source
split, and -- with world_size 2 over gloo -- the padded all-gather exchange, the column remap
and the autograd wrapper of hcspmm.dist, with the CPU oracle injected as the SpMM operator
(the product default is the CUDA path; the oracle is test infrastructure)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from helpers import rel_fro, small_graphs
from hcspmm import partition

GRAPHS = small_graphs()


def _t(a):
    return torch.from_numpy(a)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("name", ["rmat_1000", "rmat_hub_4096", "holes_777", "empty_48"])
def test_window_cuts(name, world):
    rp, ci = GRAPHS[name]
    n = rp.size - 1
    cuts = partition.window_cuts(_t(rp), world)
    assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == n
    assert all(a <= b for a, b in zip(cuts, cuts[1:]))
    assert all(c % 16 == 0 for c in cuts[:-1])
    if ci.size and world <= 4:
        per = [int(rp[cuts[i + 1]] - rp[cuts[i]]) for i in range(world)]
        wmax = max(int(rp[min(16 * (w + 1), n)] - rp[16 * w]) for w in range((n + 15) // 16))
        assert max(per) <= ci.size / world + wmax           # balanced up to one window


def test_local_shard_and_source_split_reassemble():
    rp, ci = GRAPHS["rmat_1000"]
    cuts = partition.window_cuts(_t(rp), 3)
    x = np.random.default_rng(0).standard_normal((1000, 8)).astype(np.float32)
    full = oracle.spmm(rp, ci, x, precision=1)
    for r in range(3):
        rp_l, ci_l = partition.local_shard(_t(rp), _t(ci), cuts[r], cuts[r + 1])
        y = oracle.spmm(rp_l.numpy(), ci_l.numpy(), x, precision=1)
        assert np.array_equal(y, full[cuts[r]:cuts[r + 1]])
        acc = np.zeros_like(y)
        total = 0
        for rp_s, ci_s in partition.split_by_source(rp_l, ci_l, cuts):
            total += ci_s.numel()
            acc = oracle.spmm(rp_s.numpy(), ci_s.numpy(), x, precision=1, y_init=acc)
        assert total == ci_l.numel()
        assert rel_fro(acc, y) <= 1e-6


def _oracle_ops():
    def run(x, rowptr, colidx, pre, out=None):
        y = torch.from_numpy(oracle.spmm(rowptr.numpy(), colidx.numpy(), x.detach().numpy(), precision=1))
        if out is not None:
            out.copy_(y)
            return out
        return y

    def prep(colidx, rowptr):
        return tuple(torch.from_numpy(a) for a in oracle.preprocess(colidx.numpy(), rowptr.numpy(), 0))

    return run, prep


def _worker(rank, world, port, name, out_dir, schedule="gather"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hcspmm import dist as hd
        rp, ci = GRAPHS[name]
        n = rp.size - 1
        run, prep = _oracle_ops()
        g = hd.ShardedGraph(_t(rp), _t(ci), spmm=run, preprocess=prep, schedule=schedule)
        if schedule == "halo":
            assert g.halo is not None and g.halo["rows"] <= n and sum(g.halo["recv"]) == g.halo["rows"]
        x = torch.from_numpy(np.random.default_rng(1).standard_normal((n, 12)).astype(np.float32))
        y = g.aggregate(g.shard_rows(x))
        full = oracle.spmm(rp, ci, x.numpy(), precision=1)
        assert rel_fro(y.numpy(), full[g.r0:g.r1]) <= 1e-6
        # preprocessing of the remapped shard == the slice of the global preprocessing
        gbp, getc, _, ght = oracle.preprocess(ci, rp, 0)
        e0, e1 = rp[g.r0], rp[g.r1]
        assert np.array_equal(g.pre[0].numpy(), gbp[g.r0 // 16:(g.r1 + 15) // 16])
        assert np.array_equal(g.pre[1].numpy(), getc[e0:e1])
        # a 2-layer GCN step: loss and weight gradients equal the single-process model's
        # holes_777 is NOT symmetric: backward must aggregate with (A^T)_r, a second sharded graph
        import scipy.sparse as sp
        at = sp.csr_matrix((np.ones(ci.size, np.float32), ci, rp), shape=(n, n)).T.tocsr()
        at.sort_indices()
        symmetric = (at.indptr == rp).all() and (at.indices == ci).all()
        gt = None if symmetric else hd.ShardedGraph(_t(at.indptr.astype(np.int32)), _t(at.indices.astype(np.int32)),
                                                    spmm=run, preprocess=prep, cuts=g.cuts, schedule=schedule)
        assert symmetric == (name != "holes_777")
        model = hd.DistGCN(g, 12, 8, 4, num_layers=2, seed=3, graph_t=gt)
        labels = torch.from_numpy(np.random.default_rng(2).integers(0, 4, n))
        loss = model.loss(g.shard_rows(x), labels[g.r0:g.r1])
        loss.backward()
        model.sync_grads()
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        a = torch.sparse_csr_tensor(_t(rp.astype(np.int64)), _t(ci.astype(np.int64)), torch.ones(ci.size), size=(n, n))
        ws = [w.detach().clone().requires_grad_(True) for w in model.weights]
        h = model.pad_features(x)                 # widths are laid out as multiples of 8 (zero columns / zero weights)
        assert h.shape[1] == 16 and ws[0].shape == (16, 8) and ws[1].shape == (8, 8)
        for i, w in enumerate(ws):
            h = torch.sparse.mm(a, h @ w)
            if i + 1 < len(ws):
                h = torch.relu(h)
        ref = torch.nn.functional.nll_loss(torch.log_softmax(h[:, :model.classes], 1), labels)
        ref.backward()
        assert abs(float(lt) - float(ref.detach())) <= 1e-4 * max(1.0, abs(float(ref.detach())))
        for w, p in zip(ws, model.weights):
            assert rel_fro(p.grad.numpy(), w.grad.numpy()) <= 1e-4
        assert not model.weights[0].grad[12:].any() and not model.weights[1].grad[:, 4:].any()   # padding stays untrained
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("schedule", ["gather", "halo", "auto"])
@pytest.mark.parametrize("name", ["rmat_1000", "holes_777"])
def test_sharded_aggregate_and_gcn_step_gloo_world2(tmp_path, name, schedule):
    port = 29500 + (os.getpid() % 500) + (7 if name == "holes_777" else 0) + {"gather": 0, "halo": 13, "auto": 29}[schedule]
    mp.spawn(_worker, args=(2, port, name, str(tmp_path), schedule), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_halo_exchange_world3(tmp_path):
    """Three ranks: halo rows arrive from owners on both sides of the own block."""
    mp.spawn(_worker, args=(3, 29500 + (os.getpid() % 500) + 41, "rmat_1000", str(tmp_path), "halo"), nprocs=3, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(3))


@pytest.mark.parametrize("name,world,max_refs", [("rmat_hub_4096", 2, 2), ("rmat_hub_4096", 8, 1), ("rmat_1000", 3, 2),
                                                 ("holes_777", 4, 10 ** 6), ("sbm_1024", 5, 0)])
def test_segment_tags_address_the_right_rows(name, world, max_refs):
    """Segment mode of the peer exchange (hcspmm.partition.pulled_layout / tag_segments): every rank's operand keeps its
    own rows and the remote rows it references more than max_refs times; every entry's tagged id must lead -- through
    the operand layout of the rank that bits 29..31 name -- back to the entry's global column id, and the aggregation
    computed through the tags must equal the oracle's."""
    rp, ci = small_graphs()[name]
    n = rp.size - 1
    rp_t, ci_t = _t(rp), _t(ci)
    cuts = partition.window_cuts(rp_t, world)
    bounds = torch.tensor(cuts, dtype=torch.int64)
    shards, layouts, firsts = [], [], []
    for r in range(world):
        rp_l, ci_l = partition.local_shard(rp_t, ci_t, cuts[r], cuts[r + 1])
        ids, (cold, refs) = partition.pulled_layout(ci_l.to(torch.int64), bounds, r, max_refs)
        assert bool((refs <= max_refs).all())
        own = torch.bucketize(ids, bounds[1:-1], right=True)
        firsts.append(int((own < r).sum()))
        assert torch.equal(ids[firsts[r]: firsts[r] + cuts[r + 1] - cuts[r]], torch.arange(cuts[r], cuts[r + 1]))
        shards.append((rp_l, ci_l))
        layouts.append(ids)
    x = np.random.default_rng(3).standard_normal((n, 8)).astype(np.float32)
    want = oracle.spmm(rp, ci, x, precision=1)
    for r in range(world):
        rp_l, ci_l = shards[r]
        tags = partition.tag_segments(ci_l.to(torch.int64), layouts[r], bounds, r, world, firsts).to(torch.int64) & 0xffffffff
        seg, row = tags >> partition.SEG_SHIFT, tags & partition.SEG_MASK
        gid = torch.empty_like(row)
        for s_ in range(world):
            m = seg == s_
            gid[m] = layouts[(r + s_) % world][row[m]]
        assert torch.equal(gid, ci_l.to(torch.int64))
        if max_refs == 0:
            assert bool((seg == 0).all())
        # remote rows that stay at their owner are exactly the rarely referenced ones
        remote = (ci_l < cuts[r]) | (ci_l >= cuts[r + 1])
        u, c = torch.unique(ci_l[remote].to(torch.int64), return_counts=True)
        far_ids = torch.unique(ci_l.to(torch.int64)[seg != 0])
        assert torch.equal(far_ids, u[c <= max_refs])
        # the aggregation through the tags: operand of every rank = its layout's rows of X
        ops = [x[l.numpy()] for l in layouts]
        y = np.zeros((cuts[r + 1] - cuts[r], 8), np.float32)
        rows = np.repeat(np.arange(rp_l.numel() - 1), np.diff(rp_l.numpy()))
        vals = np.stack([ops[(r + int(s_)) % world][int(w_)] for s_, w_ in zip(seg.tolist(), row.tolist())]) if row.numel() else np.zeros((0, 8), np.float32)
        np.add.at(y, rows, vals)
        assert rel_fro(y, want[cuts[r]:cuts[r + 1]]) <= 1e-5


@pytest.mark.parametrize("name,blocks", [("rmat_hub_4096", 4), ("rmat_1000", 3), ("holes_777", 8), ("sbm_1024", 2)])
def test_row_block_order(name, blocks):
    """Row-block pipeline (hcspmm.partition.row_block_order): blocks are contiguous, 16-row aligned and cover the shard;
    every operand row's `first` is exactly the first block whose entries reference it, so a block never reads a row
    that arrives with a later part of the halo; entries that do not address the local operand are ignored."""
    rp, ci = small_graphs()[name]
    n = rp.size - 1
    rp_t, ci_t = _t(rp), _t(ci).clone()
    ci_t[::7] = -5                                        # entries read in place from a peer (segment-tagged: negative)
    cuts, first = partition.row_block_order(rp_t, ci_t, n, blocks)
    assert cuts[0] == 0 and cuts[-1] == n and all(a <= b for a, b in zip(cuts, cuts[1:]))
    assert all(c % 16 == 0 for c in cuts[:-1])
    want = np.full(n, blocks, np.int64)
    for b in range(blocks):
        e0, e1 = rp[cuts[b]], rp[cuts[b + 1]]
        cols = ci_t[e0:e1].numpy()
        cols = cols[cols >= 0]
        want[cols] = np.minimum(want[cols], b)
    assert np.array_equal(first.numpy(), want)
