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
    b, buf, layout = concat_block_diagonal(gs)
    assert all(o % 16 == 0 for o, _, _ in layout.values()) and buf.dtype == np.uint8
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


def test_split_runs_reconstructs_the_operator():
    """CSR part + range entries of the run-length split == the original transposed operator."""
    from hgnn_b200.sparse_ops import split_runs
    gen = torch.Generator().manual_seed(5)
    n = 120
    up = (torch.rand(n, n, generator=gen) < 0.08).float().triu(1)
    up[0, 1:9] = 1.0                                  # node 0 has many forward edges -> heavy rows
    ops = GraphOps.from_dense((up + up.t()).numpy())
    M = ops.M

    def dense(rp, col, val):
        D = np.zeros((M, M), np.float32)
        D[np.repeat(np.arange(M), np.diff(rp)), col] = val
        return D
    full = dense(ops.bt_rowptr, ops.bt_col, ops.bt_val)
    part = dense(ops.bts_rowptr, ops.bts_col, ops.bts_val)
    assert ops.bts_rng_id.size > 0 and ops.bts_col.size < 0.5 * ops.bt_col.size
    rows = np.repeat(np.arange(M), np.diff(ops.bts_rng_rowptr))
    for r, rid, v in zip(rows, ops.bts_rng_id, ops.bts_rng_val):
        part[r, ops.bts_rng_lo[rid]:ops.bts_rng_hi[rid]] += v
    assert np.array_equal(full, part)
    # generic check on a hand-made matrix with two runs in one row and a short run that must stay
    rp = np.array([0, 70, 75, 140], np.int32)
    col = np.concatenate([np.arange(5, 75), np.arange(3, 8), np.arange(0, 65)]).astype(np.int32)
    val = np.concatenate([np.ones(70), 2 * np.ones(5), 3 * np.ones(65)]).astype(np.float32)
    rp2, c2, v2, rrp, rid, rv, lo, hi = split_runs(3, rp, col, val)
    assert rp2.tolist() == [0, 0, 5, 5] and rrp.tolist() == [0, 1, 1, 2]
    assert sorted(zip(lo.tolist(), hi.tolist())) == [(0, 65), (5, 75)] and rv.tolist() == [1.0, 3.0]


def test_deferred_offsets_equal_host_offsets():
    """The raw copy + fix-up table (applied on the GPU by hgnn_fixup_offsets; here by its numpy twin)
    gives exactly the arrays of the host-side offset path."""
    from hgnn_b200.sparse_ops import apply_fixups
    gen = torch.Generator().manual_seed(9)
    gs = []
    for n in (90, 5, 140, 1, 64):
        up = (torch.rand(n, n, generator=gen) < 0.1).float().triu(1)
        if n > 20:
            up[0, 1:12] = 1.0
        gs.append(GraphOps.from_dense((up + up.t()).numpy()))
    ref, _, _ = concat_block_diagonal(gs)
    got, buf, layout = concat_block_diagonal(gs, defer_offsets=True)
    apply_fixups(buf, layout, len(gs))
    for k, v in ref.items():
        assert np.array_equal(got[k], v), k


def test_collapsed_line_graph_invariants():
    """sparse_ops.GraphOps._build_collapsed: the phantom rows E+1..M-1 of the reference's line graph
    (functions/operators.py:59,68-71) are identical copies that no operator reads from, so one representative with
    weight M-E-1 stands for them.  Checked on the dense operators: identical rows, empty columns, weights summing
    to M, and btc = (diag(ew+) AL restricted to the active rows)^T."""
    import numpy as np
    from hgnn_b200.sparse_ops import GraphOps
    rng = np.random.default_rng(3)
    for n, p, weighted in ((12, 0.3, False), (9, 0.5, True), (5, 0.9, False), (3, 1.0, False), (4, 0.0, False)):
        up = np.triu((rng.random((n, n)) < p).astype(np.float32), 1)
        if weighted:
            up *= rng.choice([1.0, 1.5, 2.0, 3.0], size=(n, n)).astype(np.float32)
        A = up + up.T
        g = GraphOps.from_dense(A, dual=True)
        M, E = g.M, g.E
        if M == 0:
            assert g.erow.shape[0] == 0 and g.ew.shape[0] == 0
            continue
        W, WL, Pm, Pd = g.dense()
        AL = WL[:, :, 2]
        wpos = np.maximum(g.ew, 0.0)
        assert wpos.sum() == M
        act = g.erow.astype(np.int64)
        assert np.array_equal(act, np.arange(act.shape[0]))
        skipped = np.nonzero(g.ew <= 0)[0]
        if skipped.size:
            rep = E + 1
            assert g.ew[rep] == M - E - 1 and np.array_equal(skipped, np.arange(E + 2, M))
            assert np.array_equal(skipped + g.ew[skipped].astype(np.int64), np.full(skipped.shape, rep))
            for r in skipped:      # identical copies of the representative, read by nobody
                assert np.array_equal(AL[r], AL[rep]) and g.dl[r] == g.dl[rep]
                assert not AL[:, r].any() and not Pm[:, r].any() and not Pd[:, r].any()
            assert not AL[:, rep].any() and not Pm[:, rep].any()
        else:
            assert (g.ew == 1).all() and act.shape[0] == M
        btc = np.zeros((M, M), np.float32)
        r = np.repeat(np.arange(M), np.diff(g.btc_rowptr))
        btc[r, g.btc_col] = g.btc_val
        want = (wpos[:, None] * AL).T
        want[:, g.ew <= 0] = 0
        assert np.array_equal(btc, want)
