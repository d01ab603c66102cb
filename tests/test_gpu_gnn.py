"""GPU tests of the caller layer (A10): hcspmm.gnn's GCN / GIN (mirror of the reference's GNN_model.py)
trained for a few steps through the HCSPMM entry points, against the same network written with
torch.sparse.mm + torch.mm on the CPU in FP32 (the north star's reference arithmetic)."""
import numpy as np
import pytest
import torch

from helpers import rel_fro, small_graphs

pytestmark = pytest.mark.gpu
GRAPHS = small_graphs()


def cpu_reference(model_kind, rp, ci, x, y, weights, steps, lr):
    n = rp.size - 1
    a = torch.sparse_csr_tensor(torch.from_numpy(rp.astype(np.int64)), torch.from_numpy(ci.astype(np.int64)),
                                torch.ones(ci.size), size=(n, n))
    ws = [torch.nn.Parameter(w.clone()) for w in weights]
    opt = torch.optim.Adam(ws, lr=lr)
    losses, first_grads = [], None
    for _ in range(steps):
        opt.zero_grad()
        h = x
        for i, w in enumerate(ws):
            h = torch.sparse.mm(a, h @ w) if model_kind == "gcn" else torch.sparse.mm(a, h) @ w
            if i + 1 < len(ws):
                h = torch.relu(h)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(h, 1), y)
        loss.backward()
        if first_grads is None:
            first_grads = [w.grad.clone() for w in ws]
        opt.step()
        losses.append(float(loss))
    return losses, first_grads


@pytest.mark.parametrize("hidden", [32, 64])
@pytest.mark.parametrize("kind", ["gcn", "gin"])
def test_training_matches_cpu_reference(kind, hidden):
    import HCSPMM
    from hcspmm import gnn
    rp, ci = GRAPHS["rmat_1000"]          # symmetric: backward may reuse A like the reference
    n, in_dim, classes, layers, steps, lr = 1000, 24, 8, 3, 4, 0.01
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, in_dim, generator=g) * 0.1
    y = torch.randint(0, classes, (n,), generator=g)
    HCSPMM.set_classifier("shipped")
    graph = gnn.prepare(torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda())
    torch.manual_seed(1)
    net = gnn.Net(graph, in_dim, hidden, classes, num_layers=layers, model=kind, dropout=False)
    for conv in [net.conv1, *net.hidden_layers, net.conv2]:
        conv.reset_parameters()           # scaled init: un-scaled randn (reference default) explodes in 3 layers
    weights = [c.weights.detach().clone() for c in [net.conv1, *net.hidden_layers, net.conv2]]
    net = net.cuda()
    xd, yd = x.cuda(), y.cuda()

    # gradient of the first step
    loss = torch.nn.functional.nll_loss(net(xd), yd)
    loss.backward()
    grads = [c.weights.grad.detach().cpu() for c in [net.conv1, *net.hidden_layers, net.conv2]]
    net.zero_grad()
    losses, _ = gnn.train_epochs(net, xd, yd, steps, lr=lr)
    ref_losses, ref_grads = cpu_reference(kind, rp, ci, x, y, weights, steps, lr)
    assert abs(float(loss) - ref_losses[0]) <= 1e-4 * abs(ref_losses[0])
    for gd, gr in zip(grads, ref_grads):
        assert rel_fro(gd.numpy(), gr.numpy()) <= 5e-3      # fused backward GEMMs are TF32 (reference: wmma TF32)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-3 * abs(b)


def test_backward_uses_transpose_for_directed_graph():
    import HCSPMM
    from hcspmm import gnn
    rp, ci = GRAPHS["holes_777"]          # not symmetric
    n = 777
    graph = gnn.prepare(torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda(), symmetric=False)
    gnn.use_transpose(graph)
    x = torch.randn(n, 16, generator=torch.Generator().manual_seed(3), requires_grad=True)
    xd = x.detach().cuda().requires_grad_(True)
    y = gnn.HCSPMMFunction_SAG.apply(xd, *graph.args())
    y.sum().backward()
    a = torch.sparse_csr_tensor(torch.from_numpy(rp.astype(np.int64)), torch.from_numpy(ci.astype(np.int64)),
                                torch.ones(ci.size), size=(n, n))
    torch.sparse.mm(a, x).sum().backward()
    assert rel_fro(xd.grad.cpu().numpy(), x.grad.numpy()) <= 1e-5


def test_sag_profile_runs():
    from hcspmm import gnn
    rp, ci = GRAPHS["ring3_256"]
    graph = gnn.prepare(torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda())
    ms = gnn.SAG(*graph.args()).profile(torch.randn(256, 32, device="cuda"), num_rounds=5, verbose=False)
    assert ms > 0
