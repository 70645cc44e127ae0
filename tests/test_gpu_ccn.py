"""GPU parity tests for the CCN contraction kernels: golden vectors from the reference (collapse6to3,
outer_contract with a general adjacency, CCN_2D / CCN_1D forward + all gradients) and the CPU oracle
on QM9-shaped graphs; tolerance 1e-4 relative, fp32."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def T(x):
    return torch.from_numpy(np.asarray(x))


def test_collapse_and_outer_contract_golden():
    import hgnn_b200  # noqa: F401
    from hgnn_b200.functions.contraction import collapse6to3
    from hgnn_b200.functions.utils_ccn import CompnetUtils
    g = load_golden("ccn")
    out = collapse6to3(T(g["collapse/F"]).cuda())
    assert rel_err(out.cpu(), g["collapse/out"]) < TOL
    util = CompnetUtils(True)
    Tt = T(g["contract/T"]).cuda().requires_grad_()
    y = util.outer_contract(Tt, T(g["contract/adj"]).cuda())
    assert rel_err(y.detach().cpu(), g["contract/out"]) < TOL
    (y * T(g["contract/gout"]).cuda()).sum().backward()
    assert rel_err(Tt.grad.cpu(), g["contract/gT"]) < TOL
    # blocks 7..15 are bitwise identical, as in the reference (contraction.py:72-80)
    o = y.detach().view(4, 4, 18, 3)
    for k in range(7, 15):
        assert torch.equal(o[:, :, 6], o[:, :, k])


@pytest.mark.parametrize("order", [2, 1])
def test_ccn_models_golden(order):
    import hgnn_b200  # noqa: F401
    from hgnn_b200.models.compnets.model_ccn import CCN_1D, CCN_2D
    g = load_golden("ccn")
    pre = "ccn%d/param/" % order
    net = (CCN_2D if order == 2 else CCN_1D)(3, 2, 2, 2, True)
    sd = {k[len(pre):]: T(v) for k, v in g.items() if k.startswith(pre)}
    assert set(sd) == set(net.state_dict())
    net.load_state_dict(sd)
    net = net.cuda()
    Xs, As, outs = [], [], []
    for gi in range(2):
        q = "ccn%d/g%d/" % (order, gi)
        X = T(g[q + "X"]).cuda().requires_grad_()
        A = T(g[q + "A"]).cuda()
        net.zero_grad()
        y = net(X, A)
        assert rel_err(y.detach().cpu(), g[q + "out"]) < TOL
        (y * T(g[q + "gout"]).cuda()).sum().backward()
        fl = 0.1 * max(float(np.abs(v).max()) for k, v in g.items() if k.startswith(q + "grad/"))
        assert rel_err(X.grad.cpu(), g[q + "gX"], fl) < TOL
        for k, v in net.named_parameters():
            assert rel_err(v.grad.cpu(), g[q + "grad/" + k], fl) < TOL, (gi, k)
        Xs.append(X.detach())
        As.append(A)
        outs.append(y.detach())
    # the batched path gives the same per-graph outputs in one launch per level
    yb = net.forward_batch(Xs, As)
    assert rel_err(yb.detach().cpu(), torch.stack(outs).cpu()) < 1e-5


@pytest.mark.parametrize("order", [2, 1])
def test_ccn_vs_oracle_qm9_shaped(order):
    import hgnn_b200  # noqa: F401
    from hgnn_b200 import synth
    from hgnn_b200.models.compnets.model_ccn import CCN_1D, CCN_2D
    from oracle import hgnn_oracle as O
    n_layers, hidden = 3, 4
    p = O.init_ccn_params(order, 5, 1, hidden, n_layers, seed=order)
    for v in p.values():
        v.requires_grad_()
    insts = [synth.qm9_shaped_instance(i, sparse=True) for i in range(6)]
    Xs = [i[0] for i in insts]
    As = [i[1] + torch.eye(i[1].shape[0]) for i in insts]          # scripts/train_ccn.py:36
    fwd = O.ccn2_forward if order == 2 else O.ccn1_forward
    oy = torch.stack([fwd(p, n_layers, x, a) for x, a in zip(Xs, As)])
    G = torch.randn(oy.shape, generator=torch.Generator().manual_seed(3))
    (oy * G).sum().backward()
    net = (CCN_2D if order == 2 else CCN_1D)(5, 1, hidden, n_layers, True)
    net.load_state_dict({k: v.detach() for k, v in p.items()})
    net = net.cuda()
    y = net.forward_batch([x.cuda() for x in Xs], As)
    assert rel_err(y.detach().cpu(), oy.detach()) < TOL
    (y * G.cuda()).sum().backward()
    fl = 0.1 * max(float(v.grad.abs().max()) for v in p.values())
    for k, v in net.named_parameters():
        assert rel_err(v.grad.cpu(), p[k].grad, fl) < TOL, k
