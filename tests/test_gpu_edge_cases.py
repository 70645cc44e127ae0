"""GPU: ragged / degenerate inputs through the model path vs the CPU oracle: graphs with a single
edge, with no edges at all (zero line-graph nodes), a batch of one graph, J=3 powers, very uneven
sizes (heavy padding in the reference layout)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _adj(n, p, gen, weighted=False):
    up = (torch.rand(n, n, generator=gen) < p).float().triu(1)
    if weighted:
        up = up * torch.tensor([1.0, 1.5, 2.0, 3.0])[torch.randint(0, 4, (n, n), generator=gen)]
    return up + up.t()


def _compare(kind, order, h, J, adjs, seed, L=3):
    import hgnn_b200  # noqa: F401
    from hgnn_b200.functions.batching import prepare_batch
    from hgnn_b200.functions.operators import graph_operators
    from hgnn_b200.models.gnns.model_mnb import GNN_lg, GNN_simple
    from oracle import hgnn_oracle as O
    gen = torch.Generator().manual_seed(seed)
    inst, oinst = [], []
    for A in adjs:
        x = torch.randn(A.shape[0], 5, generator=gen)
        t = torch.zeros(13)
        inst.append([x, A, t] + list(graph_operators([x, A], J, True, sparse=True)))
        oinst.append([x, A, t] + list(O.graph_operators([x, A], J, True)))
    p = O.init_gnn_params(kind, h, L, 5, 2, J, max(order, 1), seed=seed)
    for v in p.values():
        v.requires_grad_()
    oX, oW, _, oXL, oWL, oPm, oPd, omask, omask_lg, oN, oE = O.prepare_batch(oinst, 0, J)
    oX.requires_grad_()
    if kind == "simple":
        oy = O.gnn_simple_forward(p, L, [oX, oW], oN, omask)
        model = GNN_simple(0, h, L, 5, 2, J)
    else:
        oy = O.gnn_lg_forward(p, L, order, [oX, oXL, oW, oWL, oPm, oPd], oN, omask, oE, omask_lg)
        model = GNN_lg(0, h, L, 5, 2, J, order)
    G = torch.randn(oy.shape, generator=gen)
    (oy * G).sum().backward()
    model.load_state_dict({k: v.detach() for k, v in p.items()})
    model = model.cuda().train()
    X, W, _, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, J)
    assert torch.equal(X, oX.detach()) and torch.equal(XL, oXL)
    assert torch.equal(N_batch, oN) and torch.equal(E_batch, oE)
    X = X.cuda().requires_grad_()
    y = (model([X, W], N_batch, mask) if kind == "simple" else
         model([X, XL.cuda(), W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg))
    assert rel_err(y.detach().cpu(), oy.detach()) < TOL
    (y * G.cuda()).sum().backward()
    fl = 0.1 * max(float(v.grad.abs().max()) for v in p.values())
    assert rel_err(X.grad.cpu(), oX.grad, fl) < TOL
    for k, v in model.named_parameters():
        assert rel_err(v.grad.cpu(), p[k].grad, fl) < TOL, k


@pytest.mark.parametrize("kind,order", [("simple", 0), ("lg", 1), ("lg", 2)])
def test_graph_without_edges_and_single_edge(kind, order):
    gen = torch.Generator().manual_seed(1)
    single = torch.zeros(3, 3)
    single[1, 2] = single[2, 1] = 2.0
    adjs = [_adj(9, 0.4, gen, True), torch.zeros(4, 4), single, _adj(6, 0.5, gen)]
    _compare(kind, order, 2, 1, adjs, seed=11)


@pytest.mark.parametrize("kind,order", [("simple", 0), ("lg", 3)])
def test_batch_of_one(kind, order):
    gen = torch.Generator().manual_seed(2)
    _compare(kind, order, 2, 1, [_adj(17, 0.3, gen, True)], seed=12)


def test_very_ragged_batch_and_J3():
    gen = torch.Generator().manual_seed(3)
    adjs = [_adj(40, 0.12, gen, True), _adj(2, 1.0, gen), _adj(13, 0.3, gen), _adj(5, 0.6, gen)]
    _compare("lg", 1, 2, 3, adjs, seed=13, L=3)
    _compare("simple", 0, 4, 3, adjs, seed=14, L=4)


def test_node_zero_hub_overflows_nothing():
    """Node 0 with many forward edges: many rows of the transposed line-graph operator carry a
    run-length (phantom) range; all of them sit in one tile / CTA."""
    gen = torch.Generator().manual_seed(4)
    A = _adj(120, 0.05, gen)
    A[0, 1:100] = 1.0
    A[1:100, 0] = 1.0
    _compare("lg", 1, 2, 1, [A, _adj(30, 0.2, gen)], seed=15)


@pytest.mark.parametrize("kind,order", [("simple", 0), ("lg", 1), ("lg", 3)])
def test_wide_states_degenerate_graphs(kind, order):
    """h = 16 (32-wide states): the middle layers run on the tensor-core tile kernels (csrc/engine_wide.cuh);
    same degenerate batch as above plus a ragged one, against the CPU oracle."""
    gen = torch.Generator().manual_seed(5)
    single = torch.zeros(3, 3)
    single[1, 2] = single[2, 1] = 2.0
    adjs = [_adj(9, 0.4, gen, True), torch.zeros(4, 4), single, _adj(6, 0.5, gen), _adj(40, 0.12, gen, True)]
    _compare(kind, order, 16, 1, adjs, seed=16, L=4)
    _compare(kind, order, 16, 1, [_adj(17, 0.3, gen, True)], seed=17, L=3)


def test_wide_states_long_rows():
    """32-wide states with rows longer than the in-line gather limit (deferred to warp-cooperative gathers), more
    deferred items than the list holds, tiles whose CSR slice exceeds the staging buffers, and the run-length
    ranges of a node-0 hub."""
    gen = torch.Generator().manual_seed(6)
    hub = _adj(120, 0.05, gen)
    hub[0, 1:100] = 1.0
    hub[1:100, 0] = 1.0
    _compare("lg", 1, 16, 1, [hub, _adj(30, 0.2, gen)], seed=18)
    _compare("lg", 1, 16, 1, [_adj(50, 0.7, gen), _adj(12, 0.3, gen, True)], seed=19)
    _compare("simple", 0, 16, 2, [_adj(50, 0.7, gen), _adj(12, 0.3, gen, True)], seed=20, L=4)
