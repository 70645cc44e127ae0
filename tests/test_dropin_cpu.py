"""CPU, build container: the drop-in contract of SURVEY.md 8(b) / rows a-10, f-2, f-4.

* the reference's UNMODIFIED epoch loops (scripts/train_mnb.py:25-94, scripts/train_ccn.py:24-73) are
  imported from the read-only checkout with this package installed under the reference's module names;
  they must run - batching, handle protocol (``.requires_grad =``, ``.cuda()``), logs stand-in - up to
  the first kernel call, where a box without a GPU raises this package's "needs a CUDA device" error
  (there is no CPU fallback to fall into);
* the host glue mirrored from the reference (``get_batches``, ``three_collinear_points``) against the
  imported reference on the same seeds;
* checkpoint compatibility: a whole-module pickle and a state_dict written by the REFERENCE
  (tests/golden/ref_gnn_lg_*.pt, produced by oracle/make_golden.py) load into this package's classes.

Tests that need /root/reference are skipped where it does not exist (the GPU box)."""
import copy
import importlib.util
import io
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import reference_shim

REF = reference_shim.REFERENCE_ROOT
needs_ref = pytest.mark.skipif(not reference_shim.available(), reason="reference checkout absent")


def _import_reference_file(rel, name):
    """Import ONE file of the reference by path (its ``from functions import ...`` lines then resolve
    through sys.modules, i.e. to whatever install_aliases() registered)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def aliases():
    import hgnn_b200
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("functions", "models")}
    names = hgnn_b200.install_aliases(force=True)
    yield hgnn_b200
    for k in names:
        sys.modules.pop(k, None)
    sys.modules.update(saved)


def _sbm_instances(n_graphs=3, N=12):
    from hgnn_b200 import synth
    return synth.sbm_dataset(n_graphs, N=N, J=1)


@needs_ref
def test_reference_train_with_mnb_runs_unmodified_up_to_the_kernels(aliases):
    """scripts/train_mnb.py:25-94, imported as is.  cuda=False keeps torch's own ``.cuda()`` out of the way
    on this GPU-less box, so the loop gets through get_batches -> prepare_batch -> ``X.requires_grad = True;
    W.requires_grad = True`` (handles) -> model(...) and stops at the package's CUDA guard."""
    train_mnb = _import_reference_file("scripts/train_mnb.py", "ref_train_mnb")
    from hgnn_b200.models.gnns.model_mnb import GNN_lg, GNN_simple
    data = _sbm_instances()
    for model in (GNN_lg(0, 2, 3, 5, 2, 1, 1), GNN_simple(0, 2, 3, 5, 2, 1)):
        opt = torch.optim.Adamax(model.parameters(), lr=1e-3)
        if torch.cuda.is_available():
            model = model.cuda()
            loss, _ = train_mnb.train_with_mnb(model, data, 0, torch.nn.CrossEntropyLoss(), opt, True, 2, 0, 1)
            assert np.isfinite(loss)
        else:
            with pytest.raises(RuntimeError, match="needs a CUDA device"):
                train_mnb.train_with_mnb(model, data, 0, torch.nn.CrossEntropyLoss(), opt, False, 2, 0, 1)


@needs_ref
def test_reference_train_ccn_runs_unmodified_up_to_the_kernels(aliases):
    """scripts/train_ccn.py:24-73, imported as is (adds the self-loops, sets requires_grad on X and A)."""
    train_ccn = _import_reference_file("scripts/train_ccn.py", "ref_train_ccn")
    from hgnn_b200 import synth
    from hgnn_b200.models.compnets.model_ccn import CCN_2D
    data = [[d[0], d[1], d[2], None, None, None, None] for d in synth.qm9_shaped_dataset(2)]
    net = CCN_2D(data[0][0].shape[1], 1, 2, 2, torch.cuda.is_available())
    opt = torch.optim.Adamax(net.parameters(), lr=1e-3)
    if torch.cuda.is_available():
        net = net.cuda()
        loss, _ = train_ccn.train_ccn(net, data, 0, torch.nn.MSELoss(), opt, True, 1.0, 1.0)
        assert np.isfinite(loss)
    else:
        with pytest.raises(RuntimeError, match="CUDA"):
            train_ccn.train_ccn(net, data, 0, torch.nn.MSELoss(), opt, False, 1.0, 1.0)


@needs_ref
def test_get_batches_matches_reference():
    """functions/batching.py:52-74 - unsorted, shuffled and size-sorted batch index lists."""
    ref = reference_shim.load()
    from hgnn_b200.functions import batching
    gen = torch.Generator().manual_seed(3)
    data = [[torch.zeros(int(n), 2)] for n in torch.randint(3, 40, (23,), generator=gen)]
    for bs in (1, 4, 23, 30):
        for shuffle_batch in (False, True):
            for sort_batch in (False, True):
                random.seed(11)
                a = ref.batching.get_batches(len(data), bs, data, shuffle_batch, sort_batch)
                random.seed(11)
                b = batching.get_batches(len(data), bs, data, shuffle_batch, sort_batch)
                assert [list(map(int, x)) for x in a] == [list(map(int, x)) for x in b], (bs, shuffle_batch, sort_batch)


@needs_ref
def test_three_collinear_points_matches_reference():
    """functions/data_generator.py:45-87 on the same seeds: identical X, A, y (same draw order) and
    operators equal to the reference's dense ones (densified on the host)."""
    ref = reference_shim.load()
    import types
    # data_generator imports graph_operators from preprocessing.preprocessing (needs rdkit); that function is
    # the identical copy of functions/operators.py:11-83
    stub_pkg = types.ModuleType("preprocessing")
    stub = types.ModuleType("preprocessing.preprocessing")
    stub.graph_operators = ref.operators.graph_operators
    saved = {k: sys.modules.get(k) for k in ("preprocessing", "preprocessing.preprocessing")}
    sys.modules["preprocessing"], sys.modules["preprocessing.preprocessing"] = stub_pkg, stub
    try:
        ref_gen = _import_reference_file("functions/data_generator.py", "ref_data_generator")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    from hgnn_b200.functions import data_generator
    from hgnn_b200.functions.operators import graph_ops_of
    torch.manual_seed(5)
    random.seed(5)
    a = ref_gen.three_collinear_points(6, 9, 3, 0.5, 0.5)
    torch.manual_seed(5)
    random.seed(5)
    b = data_generator.three_collinear_points(6, 9, 3, 0.5, 0.5, sparse=True)
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1]) and torch.equal(ra[2], rb[2])
        assert ra[2].dtype == rb[2].dtype
        g = graph_ops_of(rb[1], dual=True)
        W, WL, Pm, Pd = (torch.from_numpy(v) for v in g.dense())
        assert torch.equal(W, ra[3]) and torch.equal(WL, ra[4]) and torch.equal(Pm, ra[5]) and torch.equal(Pd, ra[6])


def test_reference_pickled_module_and_state_dict_load(aliases):
    """f-2: ``torch.load`` of a gnn.pt written by the reference's Logger.save_model (functions/logs.py:99-111)
    resolves to this package's classes; the reference's plain-attribute running statistics become buffers."""
    m = torch.load(os.path.join(GOLDEN, "ref_gnn_lg_module.pt"), weights_only=False)
    from hgnn_b200.models.gnns.model_mnb import GNN_lg
    from hgnn_b200.models.layers.batch_normalization import BN
    assert type(m) is GNN_lg and m.dual and m.J == 1 and m.order == 1 and m.n_layers == 3
    z = np.load(os.path.join(GOLDEN, "checkpoint.npz"))
    for name, mod in m.named_modules():
        if isinstance(mod, BN):
            assert "running_mean" in mod._buffers and mod.n_features == 4
            assert np.array_equal(mod.running_mean.numpy(), z["running/%s.mean" % name])
            assert np.array_equal(mod.running_std.numpy(), z["running/%s.std" % name])
    sd = torch.load(os.path.join(GOLDEN, "ref_gnn_lg_state.pt"))
    fresh = GNN_lg(0, 2, 3, 5, 2, 1, 1)
    assert set(sd) == set(fresh.state_dict())
    fresh.load_state_dict(sd)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    # pickle / deepcopy round trips of the loaded module
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    m3 = copy.deepcopy(m)
    for other in (m2, m3):
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), other.state_dict().values()))
        assert torch.equal(other.layer0.bn1.running_mean, m.layer0.bn1.running_mean)


def test_plan_cache_stays_out_of_the_module_state():
    """the engine plan holds ctypes structs; it must not ride along in pickles."""
    from hgnn_b200 import engine
    from hgnn_b200.models.gnns.model_mnb import GNN_lg
    m = GNN_lg(0, 2, 4, 5, 2, 1, 1)
    plan = engine.get_plan(m)
    assert engine.get_plan(m) is plan
    assert not any("plan" in k for k in m.__dict__)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    m3 = copy.deepcopy(m)
    assert engine.get_plan(m2) is not plan and engine.get_plan(m3) is not plan
    assert engine.get_plan(m2).n_flat == plan.n_flat
