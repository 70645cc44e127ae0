"""CPU: the native batch packer (csrc/hostpack.cu) against its numpy twin
``sparse_ops.concat_block_diagonal`` - same arrays bit for bit, same result after the fix-up table is
applied, for dual / primal-only batches, ragged and empty graphs, any thread count.  Host code only:
runs without a GPU (functions/batching.py:77-185 is the reference's dense zero-padding)."""
import pickle

import numpy as np
import pytest
import torch

import hgnn_b200  # noqa: F401
from hgnn_b200 import pack, synth
from hgnn_b200.sparse_ops import GraphOps, apply_fixups, concat_block_diagonal


def _graphs(sizes, seed=0, dual=True):
    rng = np.random.default_rng(seed)
    out = []
    for n in sizes:
        A = np.triu((rng.random((n, n)) < min(1.0, 4.0 / max(n, 1))).astype(np.float32), 1)
        A = A * rng.choice(np.array([1, 1.5, 2, 3], np.float32), size=A.shape)
        out.append(GraphOps.from_dense(A + A.T, dual=dual))
    return out


def _compare(graphs, dual, skip_bt, n_threads):
    _, buf0, lay0 = concat_block_diagonal(graphs, dual=dual, skip=("bt",) if skip_bt else (), defer_offsets=True)
    buf1, lay1 = pack.host_pack(graphs, dual=dual, skip_bt=skip_bt, n_threads=n_threads)
    assert set(lay0) == set(lay1)
    for k, (o0, dt0, n0) in lay0.items():
        o1, dt1, n1 = lay1[k]
        assert (n0, dt0) == (n1, dt1), k
        if k != "fixup":     # the table holds buffer offsets, which may differ; compared after applying it
            assert np.array_equal(buf0[o0:o0 + 4 * n0].view(dt0), buf1[o1:o1 + 4 * n1].view(dt1)), k
    apply_fixups(buf0, lay0, len(graphs))
    apply_fixups(buf1, lay1, len(graphs))
    for k, (o0, dt0, n0) in lay0.items():
        if k != "fixup":
            o1 = lay1[k][0]
            assert np.array_equal(buf0[o0:o0 + 4 * n0].view(dt0), buf1[o1:o1 + 4 * n0].view(dt0)), k
    for k, (o1, _, n1) in lay1.items():
        assert o1 % 16 == 0 and o1 + 4 * n1 <= buf1.shape[0], k


@pytest.mark.parametrize("sizes", [[12, 7, 30], [5], [40, 40, 40, 40, 3, 25, 9], [6, 2, 2, 11]])
@pytest.mark.parametrize("skip_bt", [True, False])
def test_host_pack_equals_numpy_concat(sizes, skip_bt):
    graphs = _graphs(sizes, seed=len(sizes))
    for nt in (1, 3):
        _compare(graphs, True, skip_bt, nt)


def test_host_pack_primal_only_and_edgeless_graph():
    graphs = _graphs([9, 14, 4], seed=5, dual=False)
    _compare(graphs, False, False, 2)
    empty = GraphOps.from_dense(np.zeros((6, 6), np.float32), dual=True)      # no edges: M = 0
    _compare([_graphs([8], seed=1)[0], empty, _graphs([10], seed=2)[0]], True, True, 2)


def test_host_pack_sbm_batch_many_threads():
    graphs = [i[3].graph_ops for i in synth.sbm_dataset(6, N=300, J=1, sparse=True)]
    _compare(graphs, True, True, 8)


def test_blob_views_and_pickle_round_trip():
    g = _graphs([15], seed=3)[0]
    before = {f: np.array(getattr(g, f)) for f in ("a_col", "b_val", "pt_pd", "bts_rng_lo", "deg")}
    ptr = g.blob_ptr()
    assert ptr == g.blob_ptr() and ptr % 8 == 0
    for f, v in before.items():
        assert np.array_equal(getattr(g, f), v), f
    g2 = pickle.loads(pickle.dumps(g))
    assert getattr(g2, "_blob_ptr", None) is None and np.array_equal(g2.bt_col, g.bt_col) and g2.M == g.M
    _compare([g, g2], True, False, 1)


def test_host_pack_rejects_a_foreign_blob():
    import ctypes
    bad = np.zeros(64, dtype=np.int64)
    blobs = (ctypes.c_void_p * 1)(bad.ctypes.data)
    lay = (ctypes.c_longlong * (2 * hgnn_b200._lib.lib.hgnn_host_pack_n_keys()))()
    assert hgnn_b200._lib.lib.hgnn_host_pack_layout(1, blobs, 1, 0, lay) < 0
    assert b"blob" in hgnn_b200._lib.lib.hgnn_last_error()
    assert torch is not None


def test_fill_features_matches_the_python_loop():
    """hgnn_host_fill_features (the padded X / XL host tensors of prepare_batch, reference functions/batching.py:113-127,
    :171) against the per-graph slice loop, ragged graph sizes, pre-filled destination (the tails must be zeroed)."""
    import ctypes
    import numpy as np
    import torch
    from hgnn_b200 import _lib, synth
    ds = synth.sbm_dataset(3, N=50, sparse=True) + synth.sbm_dataset(2, N=37, sparse=True, first_id=7)
    graphs = [inst[3].graph_ops for inst in ds]
    bs, F = len(ds), ds[0][0].shape[1]
    Nmax, Emax = max(g.N for g in graphs), max(g.M for g in graphs)
    X, XL = torch.full((bs, F, Nmax), 7.0), torch.full((bs, 1, Emax), 7.0)
    blobs = (ctypes.c_void_p * bs)(*[g.blob_ptr() for g in graphs])
    rows = (ctypes.c_void_p * bs)(*[inst[0].data_ptr() for inst in ds])
    assert _lib.lib.hgnn_host_fill_features(bs, blobs, rows, F, Nmax, X.data_ptr(), Emax, XL.data_ptr()) == 0
    Xr, XLr = torch.zeros(bs, F, Nmax), torch.zeros(bs, 1, Emax)
    for i, inst in enumerate(ds):
        g = graphs[i]
        Xr[i, :, :g.N] = inst[0].T
        XLr[i, 0, :g.M] = torch.from_numpy(np.asarray(g.dl))
    assert torch.equal(X, Xr) and torch.equal(XL, XLr)
    # a graph wider than the padded tensor is refused
    assert _lib.lib.hgnn_host_fill_features(bs, blobs, rows, F, Nmax - 1, X.data_ptr(), Emax, XL.data_ptr()) != 0
