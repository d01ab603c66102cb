"""Host-side mirror of the reference's GNN_model.py (/root/reference/GNN_model.py:26-302): the
autograd Functions and GCNConv / GINConv / SAG modules that route Aggregation (A*X) and Update
(*W) to the `HCSPMM` entry points -- same class names, constructor arguments and call
signatures, so HC-SpMM_main.py-style scripts run unchanged.

Not a copy: the reference spells out nine near-identical Function classes; here one table
(`_ROUTES`) states, per class, which HCSPMM entry point serves the forward and the backward
aggregation and on which side of the aggregation the Update GEMM sits.  Differences from the
reference, all deliberate:
  * any hidden width (the reference's *_fixed32* routing needs hidden == 32, GNN_model.py:277-302);
  * correct gradients: `weights.transpose(0,1)` is honoured (the reference kernel reads the view's
    raw memory).  HCSPMM.set_bug_compat(True) restores the reference's reading
    for loss-curve comparison;
  * HCSPMMFunction_SAG.backward works (the reference passes a 10th argument and raises TypeError,
    GNN_model.py:54);
  * backward aggregates with A^T when a transposed graph is supplied (`Graph.t`); the reference
    reuses A, i.e. assumes a symmetric adjacency (GNN_model.py:77, paper Eq. 4).
"""
from __future__ import annotations

import time
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

import HCSPMM


@dataclass
class Graph:
    """The eight graph tensors every reference call passes positionally, plus optional A^T."""
    row_pointers: torch.Tensor
    column_index: torch.Tensor
    blockPartition: torch.Tensor
    edgeToColumn: torch.Tensor
    edgeToRow: torch.Tensor
    hybrid_type: torch.Tensor
    row_nzr: torch.Tensor
    col_nzr: torch.Tensor
    t: Optional["Graph"] = None          # transposed graph for backward (None: A symmetric)

    def args(self):
        return (self.row_pointers, self.column_index, self.blockPartition, self.edgeToColumn, self.edgeToRow,
                self.hybrid_type, self.row_nzr, self.col_nzr)

    @property
    def num_nodes(self):
        return self.row_pointers.numel() - 1


def prepare(row_pointers: torch.Tensor, column_index: torch.Tensor, symmetric: bool = True) -> Graph:
    """HCSPMM.preprocess on device CSR (HC-SpMM_main.py:52) -> Graph; with symmetric=False also
    builds and preprocesses A^T for the backward aggregation."""
    n, nnz = row_pointers.numel() - 1, column_index.numel()
    pre = HCSPMM.preprocess(column_index, row_pointers, n, nnz, (n + 15) // 16)
    g = Graph(row_pointers, column_index, *pre)
    if not symmetric:
        rows = torch.repeat_interleave(torch.arange(n, device=row_pointers.device),
                                       (row_pointers[1:] - row_pointers[:-1]).long())
        key = torch.sort(column_index.long() * n + rows).values        # (col, row) ascending
        t_cols = (key % n).to(torch.int32)
        cnt = torch.bincount(torch.div(key, n, rounding_mode="floor"), minlength=n)
        t_rp = torch.zeros(n + 1, dtype=torch.int64, device=row_pointers.device)
        t_rp[1:] = torch.cumsum(cnt, 0)
        t_rp = t_rp.to(torch.int32)
        g.t = Graph(t_rp, t_cols, *HCSPMM.preprocess(t_cols, t_rp, n, nnz, (n + 15) // 16))
        use_transpose(g)          # a directed graph must never backpropagate through A
    return g


def _graph_of(tensors) -> Graph:
    return Graph(*tensors)


# name -> (update side, forward aggregation, backward: fused entry point or plain aggregation)
#   "pre":  X' = A (X W)         (GCN, GNN_model.py:61-162)
#   "post": X' = (A X) W         (GIN, GNN_model.py:166-232)
_ROUTES = {
    "HCSPMMFunction":            ("pre", "forward", None, "forward"),
    "HCSPMMFunctionFixed32":     ("pre", "forward_fixed32", "forward_fixed32_fused", None),
    "HCSPMMFunctionFinal":       ("pre", "forward", "forward_final_fused", None),
    "HCSPMMFunctionFirst":       ("pre", "forward_fixed32", None, "forward_fixed32"),
    "HCSPMMFunction_GINFixed32": ("post", "forward_fixed32_fused", None, "forward_fixed32"),
    "HCSPMMFunction_GINFirst":   ("post", "forward", None, "forward"),
    "HCSPMMFunction_GINFinal":   ("post", "forward_GIN_final_fused", None, "forward_fixed32"),
}

# row_pointers address -> (weak reference to that row_pointers tensor, transposed Graph).  The reference's call
# signature has no slot for A^T, so the Functions find it through the forward graph's row_pointers; the weak
# reference ties an entry to the LIFETIME of that tensor: once the graph is freed the entry is dead, and an
# unrelated graph the caching allocator later places at the same address never inherits a stale A^T.
_BACKWARD_GRAPH = {}


def use_transpose(g: Graph) -> None:
    """Make every Function aggregate with g.t in backward for the graph whose row_pointers is g's
    (prepare(symmetric=False) does this itself)."""
    if g.t is not None:
        key = g.row_pointers.data_ptr()
        _BACKWARD_GRAPH[key] = (weakref.ref(g.row_pointers, lambda _r, k=key: _BACKWARD_GRAPH.pop(k, None)), g.t)


def _bwd_args(gt):
    hit = _BACKWARD_GRAPH.get(gt[0].data_ptr())
    if hit is not None:
        owner = hit[0]()
        if owner is not None and owner.data_ptr() == gt[0].data_ptr() and owner.numel() == gt[0].numel():
            return hit[1].args()
    return gt


def _make_function(name: str):
    side, fwd_name, bwd_fused_name, bwd_plain_name = _ROUTES[name]
    takes_output = name == "HCSPMMFunctionFinal"

    class _F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, X, weights, *rest):
            gt, output = (rest[:8], rest[8]) if takes_output else (rest[:8], None)
            fwd = getattr(HCSPMM, fwd_name)
            if side == "pre":                       # Update then Aggregation
                ctx.save_for_backward(X, weights, *gt, *([output] if takes_output else []))
                return fwd(torch.mm(X, weights), *gt)[0]
            if fwd_name.endswith("fused"):          # fused Aggregation + Update: [out, A X]
                out, agg = fwd(X, *gt, weights)
            else:
                agg = fwd(X, *gt)[0]
                out = torch.mm(agg, weights)
            ctx.save_for_backward(agg, weights, *gt)
            return out

        @staticmethod
        def backward(ctx, d_output):
            saved = ctx.saved_tensors
            first, weights, gt = saved[0], saved[1], saved[2:10]
            bg = _bwd_args(gt)
            d_output = d_output.contiguous()
            if side == "pre":
                X = first
                if bwd_fused_name is not None:      # dX = (A dY) W^T fused; also returns A dY
                    extra = (saved[10],) if takes_output else ()
                    d_input, d_agg = getattr(HCSPMM, bwd_fused_name)(d_output, *bg, weights.transpose(0, 1), *extra)[:2]
                else:
                    d_agg = getattr(HCSPMM, bwd_plain_name)(d_output, *bg)[0]
                    d_input = torch.mm(d_agg, weights.transpose(0, 1))
                d_weights = torch.mm(X.transpose(0, 1), d_agg)
            else:
                agg = first
                d_agg = torch.mm(d_output, weights.transpose(0, 1))
                d_weights = torch.mm(agg.transpose(0, 1), d_output)
                d_input = getattr(HCSPMM, bwd_plain_name)(d_agg.contiguous(), *bg)[0]
            return (d_input, d_weights) + (None,) * (9 if takes_output else 8)

    _F.__name__ = _F.__qualname__ = name
    return _F


HCSPMMFunction = _make_function("HCSPMMFunction")
HCSPMMFunctionFixed32 = _make_function("HCSPMMFunctionFixed32")
HCSPMMFunctionFinal = _make_function("HCSPMMFunctionFinal")
HCSPMMFunctionFirst = _make_function("HCSPMMFunctionFirst")
HCSPMMFunction_GINFixed32 = _make_function("HCSPMMFunction_GINFixed32")
HCSPMMFunction_GINFirst = _make_function("HCSPMMFunction_GINFirst")
HCSPMMFunction_GINFinal = _make_function("HCSPMMFunction_GINFinal")


class HCSPMMFunction_SAG(torch.autograd.Function):
    """Aggregation only (GNN_model.py:26-57)."""

    @staticmethod
    def forward(ctx, X, *gt):
        ctx.save_for_backward(*gt)
        return HCSPMM.forward_fixed32(X, *gt)[0]

    @staticmethod
    def backward(ctx, d_output):
        return (HCSPMM.forward(d_output.contiguous(), *_bwd_args(ctx.saved_tensors))[0],) + (None,) * 8


class SAG(torch.nn.Module):
    """GNN_model.py:236-262: holds the graph tensors; profile() is the single-kernel benchmark."""

    def __init__(self, row_pointers, column_index, blockPartition, edgeToColumn, edgeToRow, hybrid_type,
                 row_nzr, col_nzr):
        super().__init__()
        self.graph = (row_pointers, column_index, blockPartition, edgeToColumn, edgeToRow, hybrid_type,
                      row_nzr, col_nzr)

    def forward(self, X):
        return HCSPMMFunction_SAG.apply(X, *self.graph)

    def profile(self, X, num_rounds=200, verbose=True):
        """Average ms per aggregation.  The reference times a Python loop with the wall clock
        (GNN_model.py:251-261); this uses CUDA events on the launching stream."""
        for _ in range(3):
            self.forward(X)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(num_rounds):
            self.forward(X)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / num_rounds
        if verbose:
            print("=> SAG profiling avg (ms): {:.3f}".format(ms))
        return ms


class _Conv(torch.nn.Module):
    _by_fixed = {}

    def __init__(self, input_dim, output_dim, fixed=0):
        super().__init__()
        # un-scaled randn like the reference (GNN_model.py:267,290); reset_parameters() is the
        # scaled initialisation the reference defines but leaves commented out (:268-273)
        self.weights = torch.nn.Parameter(torch.randn(input_dim, output_dim))
        self.fixed = fixed

    def reset_parameters(self):
        stdv = 1.0 / (self.weights.size(1) ** 0.5)
        self.weights.data.uniform_(-stdv, stdv)

    def forward(self, X, row_pointers, column_index, blockPartition, edgeToColumn, edgeToRow, hybrid_type,
                row_nzr, col_nzr, output=None):
        gt = (row_pointers, column_index, blockPartition, edgeToColumn, edgeToRow, hybrid_type, row_nzr, col_nzr)
        fn = self._by_fixed[self.fixed]
        if fn is HCSPMMFunctionFinal:
            if output is None or output.shape != (row_pointers.numel() - 1, self.weights.size(0)):
                output = torch.empty(row_pointers.numel() - 1, self.weights.size(0), device=X.device)
            return fn.apply(X, self.weights, *gt, output)
        return fn.apply(X, self.weights, *gt)


class GCNConv(_Conv):
    """GNN_model.py:264-283: fixed = 1 first layer, 0 hidden layer, 2 output layer."""
    _by_fixed = {0: HCSPMMFunctionFixed32, 2: HCSPMMFunctionFinal, 1: HCSPMMFunctionFirst}


class GINConv(_Conv):
    """GNN_model.py:286-302."""
    _by_fixed = {0: HCSPMMFunction_GINFixed32, 2: HCSPMMFunction_GINFinal, 1: HCSPMMFunction_GINFirst}


class Net(torch.nn.Module):
    """The model HC-SpMM_main.py builds (:67-110): conv1 -> relu -> dropout -> (num_layers-2) hidden
    convs with relu -> conv2 -> log_softmax."""

    def __init__(self, graph: Graph, in_dim: int, hidden: int, classes: int, num_layers: int = 2,
                 model: str = "gcn", dropout: bool = True):
        super().__init__()
        conv = GCNConv if model == "gcn" else GINConv
        self.graph = graph
        self.conv1 = conv(in_dim, hidden, 1)
        self.hidden_layers = torch.nn.ModuleList([conv(hidden, hidden, 0) for _ in range(num_layers - 2)])
        self.conv2 = conv(hidden, classes, 2)
        self.dropout = dropout
        self.register_buffer("output", torch.zeros(graph.num_nodes, hidden), persistent=False)

    def forward(self, x):
        g = self.graph.args()
        x = torch.relu(self.conv1(x, *g, self.output))
        if self.dropout:
            x = torch.nn.functional.dropout(x, training=self.training)
        for layer in self.hidden_layers:
            x = torch.relu(layer(x, *g, self.output))
        x = self.conv2(x, *g, self.output)
        return torch.nn.functional.log_softmax(x, dim=1)


def train_epochs(model: Net, x, y, epochs: int, lr: float = 0.01, warmup: int = 0):
    """HC-SpMM_main.py:115-139 (Adam, nll_loss); returns (losses, median epoch ms by CUDA events)."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses, times = [], []
    for ep in range(warmup + epochs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model.train()
        opt.zero_grad()
        loss = torch.nn.functional.nll_loss(model(x), y)
        loss.backward()
        opt.step()
        b.record()
        torch.cuda.synchronize()
        if ep >= warmup:
            losses.append(float(loss))
            times.append(a.elapsed_time(b))
    times.sort()
    return losses, times[len(times) // 2] if times else float("nan")
