"""CPU, build container only: the oracle against the IMPORTED reference on fresh random inputs.
Skipped where the reference checkout does not exist (the GPU box)."""
import pytest
import torch

from conftest import rel_err
from oracle import hgnn_oracle as O
from oracle import reference_shim

pytestmark = pytest.mark.skipif(not reference_shim.available(), reason="reference checkout absent")


def _graph(n, p, gen, weighted):
    up = (torch.rand(n, n, generator=gen) < p).float().triu(1)
    if weighted:
        up = up * torch.tensor([1.0, 1.5, 2.0, 3.0])[torch.randint(0, 4, (n, n), generator=gen)]
    return up + up.t()


@pytest.mark.parametrize("seed", range(6))
def test_operators_random_bit_exact(seed):
    ref = reference_shim.load()
    gen = torch.Generator().manual_seed(100 + seed)
    n = int(torch.randint(2, 20, (1,), generator=gen))
    A = _graph(n, 0.3, gen, weighted=bool(seed % 2))
    V = torch.zeros(n, 1)
    J = 1 + seed % 3
    for a, b in zip(ref.operators.graph_operators([V, A], J, True),
                    O.graph_operators([V, A], J, True)):
        assert torch.equal(a, b)


def test_lgnn_random_matches_reference():
    ref = reference_shim.load()
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    inst = []
    for n in (6, 9, 4):
        A = _graph(n, 0.5, gen, True)
        x = torch.randn(n, 5, generator=gen)
        inst.append([x, A, torch.zeros(13)] + list(ref.operators.graph_operators([x, A], 1, True)))
    batch = ref.batching.prepare_batch(inst, 0, 1)
    X, W, _, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = batch
    for a, b in zip(batch, O.prepare_batch(inst, 0, 1)):
        assert torch.equal(a, b)
    for order in (1, 2, 3):
        model = ref.model_mnb.GNN_lg(0, 4, 5, 5, 2, 1, order)
        p = {k: v.detach().clone() for k, v in model.state_dict().items()}
        y_ref = model([X, XL, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
        y = O.gnn_lg_forward(p, 5, order, [X, XL, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
        assert rel_err(y, y_ref.detach()) < 1e-4
