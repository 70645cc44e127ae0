"""CPU: the static step program handed to csrc/program.cu (engine._Plan -> hgnn_program_t): tensor table,
side list in the reference's layer order (models/gnns/model_mnb.py:58-66,124-129 over
layers_mnb.py:189-225,256-290,322-358), parameter indices, and the workspace size the native executor
derives from it.  Host logic only - no kernel runs."""
import ctypes

import pytest
import torch

import hgnn_b200  # noqa: F401
from hgnn_b200 import _lib, engine
from hgnn_b200.models.gnns.model_mnb import GNN_lg, GNN_simple


def _program(plan):
    pr = _lib.ProgramT()
    pr.n_tensors, pr.tensors = len(plan.tensors), plan.c_tensors
    pr.n_sides, pr.sides = len(plan.sides), plan.c_sides
    pr.dual, pr.arena_doubles = 1 if plan.lg else 0, plan.arena_size
    return pr


@pytest.mark.parametrize("order", [1, 2, 3])
def test_lgnn_program_follows_the_reference_layer_order(order):
    L, h, J = 5, 2, 1
    model = GNN_lg(0, h, L, 5, 2, J, order)
    plan = engine._Plan(model)
    names = list(plan.tensors)
    assert names[:2] == ["X", "XL"] and len(plan.sides) == 2 * (L - 1) + 1
    params = list(model.parameters())
    for i, s in enumerate(plan.c_sides):
        py = plan.sides[i]
        assert (s.kind == 0) == (py.kind == "node")
        assert names[s.src_self] == py.src_self
        assert (s.src_cross < 0) == (py.src_cross is None) and (s.out < 0) == (i == len(plan.sides) - 1)
        assert params[s.Wa] is py.conv_a.weight and params[s.ba] is py.conv_a.bias
        if py.conv_b is not None:
            assert params[s.Wb] is py.conv_b.weight and s.Hb == py.conv_b.weight.shape[0]
    first = plan.sides[0], plan.sides[1]
    if order == 1:      # node update first, the edge update reads the NEW node state (layers_mnb.py:203-215)
        assert first[0].kind == "node" and first[1].src_cross == first[0].out
    elif order == 2:    # edge update first, the node update reads the NEW edge state (:270-282)
        assert first[0].kind == "edge" and first[1].src_cross == first[0].out
    else:               # both read the old states (:336-348)
        assert first[0].src_cross == "XL" and first[1].src_cross == "X"
    # every normalised tensor points at its BN affine by parameter index
    for t, ct in zip(plan.tensors.values(), plan.c_tensors):
        if t["bn"] is None:
            assert ct.bn_weight == -1
        else:
            assert params[ct.bn_weight] is t["bn"].weight and params[ct.bn_bias] is t["bn"].bias
    # workspace = every side output (rows x width, 32-float aligned) + the readout rows
    Rn, Rm = 1000, 4990
    want = sum(((Rm if t["rows"] == "m" else Rn) * t["F"] + 31) // 32 * 32
               for name, t in plan.tensors.items() if name not in ("X", "XL"))
    want += (Rn * plan.readout_width + 31) // 32 * 32
    pr = _program(plan)
    assert _lib.lib.hgnn_program_work_floats(ctypes.byref(pr), Rn, Rm) == want


def test_power_gnn_program_and_gradient_table():
    model = GNN_simple(0, 2, 4, 5, 1, 2)
    plan = engine._Plan(model)
    assert not plan.lg and [s.kind for s in plan.sides] == ["node"] * 4 and plan.K == 4
    assert all(s.src_cross < 0 for s in plan.c_sides)
    # the accumulator -> flat gradient table covers every parameter exactly once, in parameters() order
    n = sum(p.numel() for p in model.parameters())
    off, nb, stride, cnt = plan.table_host
    assert plan.n_flat == n == off.shape[0] == nb.shape[0]
    assert int((off + (nb - 1).clip(0) * stride + cnt - 1).max()) < plan.arena_size
    pr = _program(plan)
    assert _lib.lib.hgnn_program_work_floats(ctypes.byref(pr), 64, 0) > 0
    assert _lib.lib.hgnn_program_work_floats(None, 64, 0) == -1
    assert torch.is_tensor(plan.params[0])
