"""GPU: the device-side batch assembly (per-graph DMA + one gather kernel, csrc/hostpack.cu) must give
exactly the arrays of the host-side block-diagonal concatenation (sparse_ops.concat_block_diagonal with
the offsets applied) - bit-exact index bookkeeping, BASELINE.json north_star."""
import numpy as np
import pytest
import torch

from test_hostpack import _graphs

pytestmark = pytest.mark.gpu


def _check(graphs, dual, skip_bt):
    from hgnn_b200 import pack
    from hgnn_b200.sparse_ops import concat_block_diagonal
    arrays, _, _ = concat_block_diagonal(graphs, dual=dual, skip=("bt",) if skip_bt else ())
    views, buf, nbytes = pack.device_pack(graphs, dual=dual, skip_bt=skip_bt, device="cuda")
    torch.cuda.synchronize()
    want = {k for k in arrays if not k.startswith("_seg")}
    assert set(views) == want, set(views) ^ want
    for k in want:
        got = views[k].cpu().numpy()
        assert got.dtype == arrays[k].dtype and np.array_equal(got, arrays[k]), k
    if not skip_bt:
        assert nbytes >= sum(g._blob.nbytes for g in graphs)


@pytest.mark.parametrize("sizes", [[12, 7, 30], [5], [40, 40, 40, 40, 3, 25, 9], [300, 2, 150]])
@pytest.mark.parametrize("skip_bt", [True, False])
def test_device_pack_equals_host_concat(sizes, skip_bt):
    _check(_graphs(sizes, seed=sum(sizes)), True, skip_bt)


def test_device_pack_primal_only_and_edgeless_graph():
    from hgnn_b200.sparse_ops import GraphOps
    _check(_graphs([9, 14, 4], seed=5, dual=False), False, False)
    empty = GraphOps.from_dense(np.zeros((6, 6), np.float32), dual=True)
    _check([_graphs([8], seed=1)[0], empty, _graphs([10], seed=2)[0]], True, True)


def test_device_and_host_concat_paths_give_the_same_pack():
    from hgnn_b200 import pack, synth
    graphs = [i[3].graph_ops for i in synth.sbm_dataset(4, N=200, J=1, sparse=True)]
    a = pack.BatchPack.from_graphs(graphs, 1, True, "cuda")
    pack.HOST_CONCAT = True
    try:
        b = pack.BatchPack.from_graphs(graphs, 1, True, "cuda")
    finally:
        pack.HOST_CONCAT = False
    for name in ("node_off", "edge_off", "pad_n", "deg", "dl"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    for ca, cb in ((a.a[0], b.a[0]), (a.at[0], b.at[0]), (a.b[0], b.b[0]), (a.p, b.p), (a.pt, b.pt), (a.bts, b.bts)):
        assert torch.equal(ca.rowptr, cb.rowptr) and torch.equal(ca.col, cb.col) and torch.equal(ca.val, cb.val)
        if ca.val2 is not None:
            assert torch.equal(ca.val2, cb.val2)
    for ra, rb in zip(a.bts_ranges, b.bts_ranges):
        assert torch.equal(ra, rb)
    assert graphs[0]._blob is not None and torch.from_numpy(graphs[0]._blob).is_pinned()


def test_batch_loader_yields_what_prepare_batch_returns():
    """functions.batching.BatchLoader = prepare_batch one batch ahead on a background thread / copy stream:
    same tuples in the same order, usable on the consumer's stream without any explicit synchronisation."""
    from hgnn_b200 import synth
    from hgnn_b200.functions.batching import BatchLoader, get_batches, prepare_batch
    from hgnn_b200.models.gnns.model_mnb import GNN_lg
    data = synth.sbm_dataset(10, N=80, J=1, sparse=True)
    idx = get_batches(len(data), 4, data)
    assert [len(i) for i in idx] == [4, 4, 2]
    torch.manual_seed(1)
    model = GNN_lg(0, 2, 3, 5, 2, 1, 1).cuda().train()
    loader = BatchLoader(data, idx, 0, 1)
    assert len(loader) == 3
    n = 0
    for k, got in enumerate(loader):
        ref = prepare_batch([data[i] for i in idx[k]], 0, 1)
        for pos in (0, 2, 3, 9, 10):                      # X, T, XL, N_batch, E_batch
            assert torch.equal(got[pos], ref[pos]), pos
        pa, pb = got[1].pack, ref[1].pack
        for name in ("node_off", "edge_off", "deg", "dl"):
            assert torch.equal(getattr(pa, name), getattr(pb, name)), name
        for ca, cb in ((pa.a[0], pb.a[0]), (pa.b[0], pb.b[0]), (pa.p, pb.p), (pa.pt, pb.pt), (pa.bts, pb.bts)):
            assert torch.equal(ca.rowptr, cb.rowptr) and torch.equal(ca.col, cb.col) and torch.equal(ca.val, cb.val)
        outs = []
        for b in (got, ref):
            X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
            outs.append(model([X.cuda(), XL.cuda(), W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg).detach())
        assert torch.equal(outs[0], outs[1]) or float((outs[0] - outs[1]).abs().max()) < 1e-6
        n += 1
    assert n == 3


def test_batch_loader_propagates_errors_and_stops_early():
    from hgnn_b200 import synth
    from hgnn_b200.functions.batching import BatchLoader
    data = synth.sbm_dataset(6, N=40, J=1, sparse=True)
    it = iter(BatchLoader(data, [[0, 1], [2, 3], [4, 5]], 0, 1))
    next(it)
    it.close()                                               # consumer leaves early: the producer thread must end
    with pytest.raises(IndexError):
        for _ in BatchLoader(data, [[0, 1], [99]], 0, 1):
            pass


def test_upload_coalesces_adjacent_blobs_and_survives_any_order():
    """Graph blobs built one after the other sit back to back in the pinned slabs and go up in ONE DMA per run
    (csrc/hostpack.cu: stage_offsets); a batch that takes them in another order, repeats one, or mixes in a blob from
    elsewhere must give the same arrays as the host concatenation; eight or more scattered small blobs are copied into
    the pinned meta buffer and go up with it."""
    graphs = _graphs([20, 9, 33, 14, 27, 6], seed=11, dual=False)
    for g in graphs:
        g.blob_ptr()                                  # allocation order = list order: one contiguous run
    ptrs = [g.blob_ptr() for g in graphs]
    assert all(0 < b - a < (1 << 20) for a, b in zip(ptrs, ptrs[1:]))
    _check(graphs, False, False)                      # one run
    _check(graphs[::-1], False, False)                # descending addresses: every blob its own DMA
    _check([graphs[2], graphs[3], graphs[0], graphs[1], graphs[5]], False, False)      # two runs and a single
    _check([graphs[1], graphs[1], graphs[2]], False, False)                            # a repeated graph
    many = _graphs([7, 12, 9, 15, 5, 11, 8, 14, 6, 10, 13, 4], seed=23, dual=False)
    for g in many:
        g.blob_ptr()
    _check(many[::-1], False, False)                  # 12 scattered small blobs: staged inside the meta buffer, one copy
    _check(many[::2] + many[1::2], False, False)
    dual_many = _graphs([9, 6, 12, 8, 7, 10, 5, 11, 9], seed=29)
    for g in dual_many:
        g.blob_ptr()
    _check(dual_many, True, True)                     # skipped tails: nine separate prefixes -> meta staging too
    _check(dual_many[::-1], True, False)
    dual = _graphs([12, 12, 30], seed=3)
    for g in dual:
        g.blob_ptr()
    _check(dual, True, True)                          # skipped transposed-operator tails: gaps too wide to join
    _check(dual, True, False)
