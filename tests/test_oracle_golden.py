"""CPU: pin the oracle (oracle/hgnn_oracle.py) against golden vectors produced by the reference
itself (oracle/make_golden.py).  Operator construction is bit-exact; activations / gradients are
within 1e-4 relative (BASELINE.json north_star tolerance, fp32)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, grad_floor
from oracle import hgnn_oracle as O

TOL = 1e-4


def T(x):
    return torch.from_numpy(np.asarray(x))


def _instances(g, J):
    inst = []
    for i in range(int(g["n_inst"])):
        x, A = T(g["inst%d/x" % i]), T(g["inst%d/A" % i])
        t = T(g["inst%d/t" % i]) if ("inst%d/t" % i) in g else torch.zeros(13)
        W, WL, Pm, Pd = O.graph_operators([x, A], J, True)
        inst.append([x, A, t, W, WL, Pm, Pd])
    return inst


def test_graph_operators_bit_exact():
    g = load_golden("operators")
    names = sorted({k.split("/")[0] for k in g})
    assert len(names) == 7
    for name in names:
        A = T(g[name + "/A"])
        V = torch.zeros(A.shape[0], 2)
        for J in (1, 2, 3):
            W, WL, Pm, Pd = O.graph_operators([V, A], J, True)
            assert torch.equal(W, T(g["%s/J%d/W" % (name, J)])), (name, J)
            assert torch.equal(WL, T(g["%s/J%d/WL" % (name, J)])), (name, J)
            assert torch.equal(Pm, T(g[name + "/Pm"])) and torch.equal(Pd, T(g[name + "/Pd"]))
            assert torch.equal(O.graph_operators([V, A], J, False), W)


def test_prepare_batch_bit_exact():
    g = load_golden("prepare_batch")
    res = O.prepare_batch(_instances(g, 2), 4, 2)
    names = ["X", "W", "T", "XL", "WL", "Pm", "Pd", "mask", "mask_lg", "N_batch", "E_batch"]
    for n, v in zip(names, res):
        assert torch.equal(v, T(g["out/" + n])), n


def test_standalone_ops():
    g = load_golden("ops")
    X, W, _, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = O.prepare_batch(_instances(g, 2), 0, 2)
    Xn, Xe = T(g["Xn"]).requires_grad_(), T(g["Xe"]).requires_grad_()
    outs = [O.graph_op(W, Xn), O.graph_op(WL, Xe), O.pmul(Pm, Xe), O.pmul(Pd, Xe),
            O.pmul(Pm.transpose(2, 1), Xn), O.pmul(Pd.transpose(2, 1), Xn)]
    for i, y in enumerate(outs):
        assert rel_err(y, g["y%d" % (i + 1)]) < TOL
    sum((y * T(g["g%d" % (i + 1)])).sum() for i, y in enumerate(outs)).backward()
    assert rel_err(Xn.grad, g["gXn"]) < TOL and rel_err(Xe.grad, g["gXe"]) < TOL
    H = T(g["bn/H"]).requires_grad_()
    w, b = T(g["bn/weight"]).requires_grad_(), T(g["bn/bias"]).requires_grad_()
    y, mean, std = O.bn_forward(H, N_batch, mask, w, b)
    assert rel_err(y, g["bn/out"]) < TOL
    (y * T(g["bn/gout"])).sum().backward()
    assert rel_err(H.grad, g["bn/gH"]) < TOL
    assert rel_err(w.grad, g["bn/gweight"]) < TOL and rel_err(b.grad, g["bn/gbias"]) < TOL
    assert rel_err(0.9 * mean, g["bn/running_mean"]) < TOL
    assert rel_err(0.9 * std, g["bn/running_std"]) < TOL


MODELS = ["gnn_simple_h3_L4_J2", "gnn_simple_h2_L3_J1", "gnn_lg1_h2_L3_J1", "gnn_lg2_h2_L3_J1",
          "gnn_lg3_h2_L3_J1", "gnn_lg1_h3_L4_J2"]


@pytest.mark.parametrize("name", MODELS)
def test_model_forward_backward(name):
    g = load_golden(name)
    J, L, order = int(g["J"]), int(g["L"]), int(g["order"])
    X, W, _, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = O.prepare_batch(_instances(g, J), 0, J)
    p = {k[len("param/"):]: T(v).requires_grad_() for k, v in g.items() if k.startswith("param/")}
    X.requires_grad_()
    if order == 0:
        l0 = O.layer_simple(p, "layer0.", [X, W], N_batch, mask)
        y = O.gnn_simple_forward(p, L, [X, W], N_batch, mask)
    else:
        state = [X, XL, W, WL, Pm, Pd]
        l0 = O.layer_with_lg(order, p, "layer0.", state, N_batch, mask, E_batch, mask_lg)
        assert rel_err(l0[1], g["layer0/XL"]) < TOL
        y = O.gnn_lg_forward(p, L, order, state, N_batch, mask, E_batch, mask_lg)
    assert rel_err(l0[0], g["layer0/X"]) < TOL
    assert rel_err(y, g["out"]) < TOL
    (y * T(g["gout"])).sum().backward()
    fl = grad_floor(g)
    assert rel_err(X.grad, g["grad/X"], fl) < TOL
    for k, v in p.items():
        assert rel_err(v.grad, g["grad/" + k], fl) < TOL, k
    # eval pass with the recorded running statistics (batch_normalization.py:39-41)
    run = {}
    for k in g:
        if k.startswith("running/") and k.endswith(".mean"):
            nm = k[len("running/"):-len(".mean")]
            run[nm] = (T(g[k]), T(g["running/%s.std" % nm]))
    stats = {"eval": run}
    with torch.no_grad():
        ye = (O.gnn_simple_forward(p, L, [X, W], N_batch, mask, stats) if order == 0 else
              O.gnn_lg_forward(p, L, order, [X, XL, W, WL, Pm, Pd], N_batch, mask, E_batch,
                               mask_lg, stats))
    assert rel_err(ye, g["out_eval"]) < TOL


def test_ccn_contraction_and_models():
    g = load_golden("ccn")
    assert rel_err(O.collapse6to3(T(g["collapse/F"])), g["collapse/out"]) < TOL
    Tt, adj = T(g["contract/T"]).requires_grad_(), T(g["contract/adj"])
    y = O.outer_contract(Tt, adj)
    assert rel_err(y, g["contract/out"]) < TOL
    assert rel_err(O.outer_contract_closed_form(Tt, adj), g["contract/out"]) < TOL
    (y * T(g["contract/gout"])).sum().backward()
    assert rel_err(Tt.grad, g["contract/gT"]) < TOL
    # the 9 repeated blocks are bitwise identical in the reference (contraction.py:72-80)
    out = T(g["contract/out"]).view(4, 4, 18, 3)
    for k in range(7, 15):
        assert torch.equal(out[:, :, 6], out[:, :, k])
    for order, fwd in ((2, O.ccn2_forward), (1, O.ccn1_forward)):
        pre = "ccn%d/param/" % order
        for gi in range(2):
            p = {k[len(pre):]: T(v).requires_grad_() for k, v in g.items() if k.startswith(pre)}
            q = "ccn%d/g%d/" % (order, gi)
            X = T(g[q + "X"]).requires_grad_()
            yo = fwd(p, 2, X, T(g[q + "A"]))
            assert rel_err(yo, g[q + "out"]) < TOL
            (yo * T(g[q + "gout"])).sum().backward()
            assert rel_err(X.grad, g[q + "gX"]) < TOL
            for k, v in p.items():
                assert rel_err(v.grad, g[q + "grad/" + k]) < TOL, (order, gi, k)


def test_workloads_match_synth():
    """oracle/workloads.py (used by the CPU baseline / reference arm of bench.py without loading the CUDA
    library) generates exactly the graphs of hgnn_b200.synth."""
    from hgnn_b200 import synth
    from oracle import workloads
    for gid, N, a, b in ((0, 40, 7.0, 3.0), (3, 61, 8.0, 2.0)):
        s = synth.sbm_instance(gid, N=N, a=a, b=b, J=1, sparse=True)
        o = workloads.sbm_dense(gid, N=N, a=a, b=b)
        assert torch.equal(s[0], o[0]) and torch.equal(s[1].to_dense(), o[1]) and torch.equal(s[2], o[2])
    for gid in (0, 5, 11):
        s = synth.qm9_shaped_instance(gid, sparse=True)
        o = workloads.qm9_shaped_dense(gid)
        assert torch.equal(s[0], o[0]) and torch.equal(s[1], o[1]) and torch.equal(s[2], o[2])
