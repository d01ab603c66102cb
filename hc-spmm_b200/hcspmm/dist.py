"""Multi-GPU aggregation: A partitioned by row windows, X exchanged per layer.

One process per GPU (torchrun), `torch.distributed` for the plumbing.  The reference has no
multi-GPU code at all; this layer sits above the HCSPMM entry points and changes nothing below.

Rank r owns the 16-row windows [cuts[r], cuts[r+1]) of A (nnz-balanced, hcspmm.partition), the
matching rows of X / Y, and the preprocessing of its shard.  One aggregation Y_r = A_r * X is
    exchange:  all-gather of the row shards of X   (NCCL over NVLink / NVSwitch)
    compute:   local hybrid SpMM on the gathered X (rectangular: n_r x N)
Two schedules:
  * "gather":   one all_gather_into_tensor, then one SpMM launch;
  * "slabs":    X is exchanged in feature slabs; the SpMM of slab k (a strided view of the gathered
                buffer, the kernel takes ldx/ldy) runs on the compute stream while slab k+1 is in
                flight on the communication stream -- the exchange hides behind the gather-bound
                kernel instead of preceding it.
Shards are padded to the largest shard so the collective is a plain equal-size all-gather; shard
s occupies rows [s*max_rows, s*max_rows + n_s) of the gathered buffer and the local column ids are
remapped to that layout once.  The remap is monotone, so the window-local column ranks
(edgeToColumn), block counts and labels are exactly those of the unpartitioned graph.

The aggregation operator is injectable (`spmm=`) so that the exchange / remap / autograd logic is
testable on CPU with gloo; the default is the CUDA path (HCSPMM), which has no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import partition


def _default_spmm():
    import HCSPMM

    def run(x, rowptr, colidx, pre, out=None):
        if out is None:
            return HCSPMM.forward(x, rowptr, colidx, *pre)[0]
        return HCSPMM.spmm_strided(x, rowptr, colidx, *pre[:4], out, False)

    def prep(colidx, rowptr):
        n = rowptr.numel() - 1
        return HCSPMM.preprocess(colidx, rowptr, n, colidx.numel(), (n + 15) // 16)

    return run, prep


class ShardedGraph:
    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, group=None, schedule: str = "gather",
                 n_slabs: int = 2, spmm=None, preprocess=None, cuts=None):
        """rowptr / colidx: the FULL graph's CSR on this rank's device (identical on every rank).
        cuts: reuse another ShardedGraph's row cuts (the transposed graph for backward must be
        partitioned like the forward one)."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n = rowptr.numel() - 1
        self.cuts = list(cuts) if cuts is not None else partition.window_cuts(rowptr, self.world)
        self.r0, self.r1 = self.cuts[self.rank], self.cuts[self.rank + 1]
        self.n_local = self.r1 - self.r0
        self.max_rows = max(self.cuts[i + 1] - self.cuts[i] for i in range(self.world))
        self.schedule, self.n_slabs = schedule, max(1, n_slabs)
        rp_l, ci_l = partition.local_shard(rowptr, colidx, self.r0, self.r1)
        dev = colidx.device
        bounds = torch.tensor(self.cuts, device=dev, dtype=torch.int64)
        owner = torch.bucketize(ci_l.to(torch.int64), bounds[1:-1], right=True)
        self.colidx = (ci_l.to(torch.int64) - bounds[owner] + owner * self.max_rows).to(torch.int32).contiguous()
        self.rowptr = rp_l
        self.nnz_local = ci_l.numel()
        if spmm is None:
            spmm, preprocess = _default_spmm()
        self._spmm = spmm
        self.pre = preprocess(self.colidx, self.rowptr) if preprocess is not None else ()
        self._bufs = {}
        self._comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    # ------------------------------------------------------------------------------------------
    def shard_rows(self, x_full: torch.Tensor) -> torch.Tensor:
        return x_full[self.r0:self.r1].contiguous()

    def _buffers(self, dim, device, dtype):
        key = (dim, device, dtype)
        if key not in self._bufs:
            self._bufs[key] = (torch.zeros(self.max_rows, dim, device=device, dtype=dtype),
                               torch.empty(self.world * self.max_rows, dim, device=device, dtype=dtype))
        return self._bufs[key]

    def aggregate(self, x_local: torch.Tensor) -> torch.Tensor:
        """Y_r = A_r * X from the row shard X_r of every rank."""
        assert x_local.shape[0] == self.n_local
        dim = x_local.shape[1]
        if self.world == 1:
            return self._spmm(x_local.contiguous(), self.rowptr, self.colidx, self.pre)
        if self.schedule == "slabs" and self.n_slabs > 1 and dim >= 8 * self.n_slabs and self._comm_stream is not None:
            return self._aggregate_slabs(x_local)
        pad, gathered = self._buffers(dim, x_local.device, x_local.dtype)
        pad[: self.n_local].copy_(x_local)
        dist.all_gather_into_tensor(gathered, pad, group=self.group)
        return self._spmm(gathered, self.rowptr, self.colidx, self.pre)

    def _aggregate_slabs(self, x_local: torch.Tensor) -> torch.Tensor:
        dim = x_local.shape[1]
        step = (dim // self.n_slabs + 3) // 4 * 4
        edges = list(range(0, dim, step)) + [dim]
        dev, dt = x_local.device, x_local.dtype
        y = torch.empty(self.n_local, dim, device=dev, dtype=dt)
        cur = torch.cuda.current_stream(dev)
        comm = self._comm_stream
        comm.wait_stream(cur)                       # x_local is produced on the current stream
        events, gathered = [], []
        for k in range(len(edges) - 1):
            w = edges[k + 1] - edges[k]
            key = ("slab", k, w, dev, dt)
            if key not in self._bufs:
                self._bufs[key] = (torch.zeros(self.max_rows, w, device=dev, dtype=dt),
                                   torch.empty(self.world * self.max_rows, w, device=dev, dtype=dt))
            pad, gat = self._bufs[key]
            with torch.cuda.stream(comm):
                pad[: self.n_local].copy_(x_local[:, edges[k]:edges[k + 1]])
                dist.all_gather_into_tensor(gat, pad, group=self.group)
                ev = torch.cuda.Event()
                ev.record(comm)
            events.append(ev)
            gathered.append(gat)
        for k, (ev, gat) in enumerate(zip(events, gathered)):
            cur.wait_event(ev)                      # slab k has landed; slab k+1.. still in flight
            self._spmm(gat, self.rowptr, self.colidx, self.pre, out=y[:, edges[k]:edges[k + 1]])
        comm.wait_stream(cur)                       # buffers are reused by the next call
        return y


class ShardedAggregate(torch.autograd.Function):
    """Autograd wrapper: forward Y_r = A_r X; backward dX_r = A_r dY (A symmetric, as the reference
    assumes -- GNN_model.py:77) or (A^T)_r dY when a transposed ShardedGraph is supplied."""

    @staticmethod
    def forward(ctx, x_local, graph: ShardedGraph, graph_t):
        ctx.graph_t = graph_t if graph_t is not None else graph
        return graph.aggregate(x_local)

    @staticmethod
    def backward(ctx, d_y):
        return ctx.graph_t.aggregate(d_y.contiguous()), None, None


class DistGCN(torch.nn.Module):
    """2+-layer GCN of HC-SpMM_main.py:67-87 on a row-partitioned graph: every rank holds its rows of
    X and of the labels and a replica of the weights; Update GEMMs are row-local, Aggregations go
    through ShardedAggregate, weight gradients are summed with one all-reduce per step."""

    def __init__(self, graph: ShardedGraph, in_dim, hidden, classes, num_layers=2, seed=0, graph_t=None,
                 order: str = "auto"):
        """order: "update_first" = A (H W) (GCN), "aggregate_first" = (A H) W (GIN), "auto" = the cheaper."""
        super().__init__()
        self.order = order
        g = torch.Generator().manual_seed(seed)            # identical replicas on every rank
        dims = [in_dim] + [hidden] * (num_layers - 1) + [classes]
        self.weights = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.randn(dims[i], dims[i + 1], generator=g) / dims[i] ** 0.5)
             for i in range(num_layers)])
        self.graph, self.graph_t = graph, graph_t

    def forward(self, x_local):
        h = x_local
        for i, w in enumerate(self.weights):
            # A (H W) = (A H) W: exchange and aggregate at the narrower of the two widths -- the
            # all-gather moves N * width * 4 bytes per rank and the SpMM gathers nnz * width * 4
            update_first = {"update_first": True, "aggregate_first": False}.get(self.order, w.shape[1] <= w.shape[0])
            if update_first:
                h = ShardedAggregate.apply(torch.mm(h, w), self.graph, self.graph_t)
            else:
                h = torch.mm(ShardedAggregate.apply(h, self.graph, self.graph_t), w)
            if i + 1 < len(self.weights):
                h = torch.relu(h)
        return torch.nn.functional.log_softmax(h, dim=1)

    def loss(self, x_local, y_local):
        """nll_loss averaged over ALL vertices (sum over local rows / N)."""
        return torch.nn.functional.nll_loss(self.forward(x_local), y_local, reduction="sum") / self.graph.n

    def sync_grads(self):
        if self.graph.world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in self.parameters()])
            dist.all_reduce(flat, group=self.graph.group)
            o = 0
            for p in self.parameters():
                p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
