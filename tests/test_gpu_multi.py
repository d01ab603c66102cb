"""Multi-GPU parity of the row-partitioned aggregation (hcspmm.dist) on REAL peers: one process per GPU
(torch.multiprocessing spawn, NCCL for the plumbing), every exchange schedule -- NVLink peer pull (default),
NCCL all-gather, NCCL halo all-to-all, feature slabs, two source passes, BF16 operand -- against the CPU oracle
and the FP32 torch.sparse result of the UNPARTITIONED graph.  Skipped on a box with fewer than two GPUs (the
world-size-2/3 gloo tests in test_partition_dist_cpu.py cover the host logic there)."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "hc-spmm_b200"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        import oracle
        from helpers import small_graphs, rel_fro, torch_sparse_ref
        from hcspmm import dist as hd
        rp, ci = small_graphs()["rmat_hub_4096"]
        n = rp.size - 1
        d_rp, d_ci = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
        results = []
        # last field: direct_refs (segment mode: remote rows referenced <= T times are read in place by the SpMM;
        # 0 = every halo row is pulled, None = auto, 10**6 = nothing is pulled at all)
        # rb = row_blocks (row-block pipeline: the halo travels in the order the shard's row blocks need it and every
        # block's SpMM starts when its part has landed; 1 = off, None = auto)
        cases = [("peer", 1, 1, "fp32", 0, 1), ("gather", 1, 1, "fp32", 0, 1), ("halo", 1, 1, "fp32", 0, 1),
                 ("peer", 2, 1, "fp32", 0, 1), ("peer", 1, 2, "fp32", 0, 1), ("peer", 1, 1, "bf16", 0, 1),
                 ("auto", 1, 1, "fp32", None, None), ("push", 1, 1, "fp32", 0, 1), ("push", 1, 1, "bf16", 0, 1),
                 ("peer", 1, 1, "fp32", 2, 1), ("peer", 1, 1, "fp32", 10 ** 6, 1), ("peer", 1, 1, "bf16", 2, 1),
                 ("peer", 1, 1, "fp32", 1, 1), ("peer", 1, 1, "fp32", 0, 4), ("peer", 1, 1, "bf16", 0, 2),
                 ("peer", 1, 1, "fp32", 2, 3), ("peer", 1, 1, "fp32", 0, 8)]
        for schedule, slabs, passes, operand, direct, rb in cases:
            g = hd.ShardedGraph(d_rp, d_ci, schedule=schedule, n_slabs=slabs, n_passes=passes, operand=operand,
                                direct_refs=direct, row_blocks=rb)
            if direct:
                assert g.direct is not None and g.direct["T"] == direct, "segment mode was not set up"
                schedule = f"{schedule}+inplace{direct}"
            if rb is not None and rb > 1:
                assert g.blocks is not None and g.blocks["B"] == rb, "row-block pipeline was not set up"
                schedule = f"{schedule}+rowblocks{rb}"
            for dim in (128, 100, 64) + ((320,) if direct else ()):
                x = np.random.default_rng(dim).standard_normal((n, dim)).astype(np.float32)
                want = oracle.spmm(rp, ci, x, precision=1)[g.r0:g.r1]
                want_ts = torch_sparse_ref(rp, ci, x)[g.r0:g.r1]
                x_loc = torch.from_numpy(x[g.r0:g.r1]).to(dev)
                for rep in range(3):                       # both operand buffers, repeatedly
                    got = g.aggregate(x_loc).cpu().numpy()
                tol = 1e-2 if operand == "bf16" else 1e-5
                e1, e2 = rel_fro(got, want), rel_fro(got, want_ts)
                results.append((schedule, slabs, passes, operand, dim, g.schedule, e1, e2, tol))
            g.check()
            g.close()
        # autograd: loss and input gradient of one aggregation layer equal the unpartitioned computation
        g = hd.ShardedGraph(d_rp, d_ci, schedule="auto")
        x = np.random.default_rng(7).standard_normal((n, 32)).astype(np.float32)
        x_loc = torch.from_numpy(x[g.r0:g.r1]).to(dev).requires_grad_(True)
        y = hd.ShardedAggregate.apply(x_loc, g, None)
        (y * y).sum().backward()
        yy = oracle.spmm(rp, ci, x, precision=1)
        grad_want = oracle.spmm(rp, ci, 2 * yy, precision=1)[g.r0:g.r1]       # A symmetric: dX = A (2 Y)
        results.append(("autograd", 1, 1, "fp32", 32, g.schedule, rel_fro(y.detach().cpu().numpy(), yy[g.r0:g.r1]),
                        rel_fro(x_loc.grad.cpu().numpy(), grad_want), 1e-5))
        g.close()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", results))
    except Exception:
        q.put((rank, "error", traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_aggregation_on_real_peers_matches_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = []
    try:
        for _ in range(world):
            out.append(q.get(timeout=600))
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    errs = [o for o in out if o[1] != "ok"]
    assert not errs, "\n".join(e[2] for e in errs)
    for rank, _, results in out:
        for (schedule, slabs, passes, operand, dim, used, e1, e2, tol) in results:
            assert e1 <= tol and e2 <= tol, (f"rank {rank} schedule {schedule}->{used} slabs {slabs} passes {passes} "
                                             f"operand {operand} dim {dim}: rel err {e1:.2e} (oracle) {e2:.2e} (torch.sparse)")
    if os.environ.get("HCSPMM_TEST_REPORT"):
        with open(os.environ["HCSPMM_TEST_REPORT"], "a") as f:
            for rank, _, results in sorted(out):
                for r in results:
                    f.write(f"world {world} rank {rank} " + " ".join(str(v) for v in r) + "\n")
