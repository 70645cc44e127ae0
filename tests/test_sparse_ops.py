"""CPU: the host-side sparse operator bookkeeping (hgnn-2_b200/sparse_ops.py) reproduces the
reference's dense operators bit-exactly (golden vectors from the reference + the oracle)."""
import numpy as np
import pytest
import torch

import hgnn_b200  # noqa: F401
from hgnn_b200.sparse_ops import GraphOps, concat_block_diagonal
from conftest import load_golden
from oracle import hgnn_oracle as O


def test_graph_ops_match_golden_operators():
    g = load_golden("operators")
    for name in sorted({k.split("/")[0] for k in g}):
        ops = GraphOps.from_dense(g[name + "/A"])
        W, WL, Pm, Pd = ops.dense()
        assert np.array_equal(W, g[name + "/J1/W"]), name
        assert np.array_equal(WL, g[name + "/J1/WL"]), name
        assert np.array_equal(Pm, g[name + "/Pm"]) and np.array_equal(Pd, g[name + "/Pd"]), name


@pytest.mark.parametrize("seed", range(8))
def test_graph_ops_random_vs_oracle(seed):
    gen = torch.Generator().manual_seed(seed)
    n = int(torch.randint(1, 40, (1,), generator=gen))
    up = (torch.rand(n, n, generator=gen) < 0.25).float().triu(1)
    if seed % 2:
        up = up * torch.tensor([1.0, 1.5, 2.0, 3.0])[torch.randint(0, 4, (n, n), generator=gen)]
    A = up + up.t()
    if seed % 3 == 0:
        A = A + torch.diag((torch.rand(n, generator=gen) < 0.3).float())
    ref = O.graph_operators([torch.zeros(n, 1), A], 1, True)
    ops = GraphOps.from_dense(A.numpy())
    for a, b in zip(ops.dense(), ref):
        assert np.array_equal(a, b.numpy())
    # transposes really are transposes
    def dn(n_r, n_c, rp, col, val):
        D = np.zeros((n_r, n_c), np.float32)
        D[np.repeat(np.arange(n_r), np.diff(rp)), col] = val
        return D
    M = ops.M
    assert np.array_equal(dn(M, M, ops.bt_rowptr, ops.bt_col, ops.bt_val), ref[1][:, :, 2].numpy().T)
    assert np.array_equal(dn(M, n, ops.pt_rowptr, ops.pt_col, ops.pt_pd), ref[3].numpy().T)
    assert np.array_equal(dn(M, n, ops.pt_rowptr, ops.pt_col, ops.pt_pm), ref[2].numpy().T)
    assert np.array_equal(dn(n, n, ops.at_rowptr, ops.at_col, ops.at_val), A.numpy().T)


def test_block_diagonal_concat():
    gen = torch.Generator().manual_seed(3)
    gs = []
    for n in (4, 7, 1, 5):
        up = (torch.rand(n, n, generator=gen) < 0.5).float().triu(1)
        gs.append(GraphOps.from_dense((up + up.t()).numpy()))
    b = concat_block_diagonal(gs)
    assert b["node_off"].tolist() == [0, 4, 11, 12, 17]
    assert b["edge_off"][-1] == sum(g.M for g in gs)
    for name, rows in (("a", 17), ("b", b["edge_off"][-1]), ("p", 17), ("pt", b["edge_off"][-1])):
        rp = b[name + "_rowptr"]
        assert rp.shape[0] == rows + 1 and rp[0] == 0 and np.all(np.diff(rp) >= 0)
    # block 1 of A sits at rows/cols 4..10
    rp, col = b["a_rowptr"], b["a_col"]
    seg = col[rp[4]:rp[11]]
    assert seg.size == gs[1].a_col.size and seg.min() >= 4 and seg.max() < 11
    assert np.array_equal(seg - 4, gs[1].a_col)
