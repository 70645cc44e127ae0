"""CPU: the C-ABI library loads and exports every symbol include/hgnn_b200.h declares (no compute
calls - there is no GPU here), and the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "hgnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hgnn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import hgnn_b200
    from hgnn_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.EXPORTS) == names, "ctypes signatures and header disagree"
    assert lib.hgnn_version() == 100
    assert hgnn_b200._lib.lib.hgnn_workspace_bytes(8) > 256


def test_library_is_sm100a_only():
    import subprocess
    from hgnn_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import hgnn_b200  # noqa: F401
    from hgnn_b200.functions.operators import graph_operators
    from hgnn_b200.models.gnns.model_mnb import GNN_simple
    from hgnn_b200.models.layers.layers_mnb import graph_oper
    A = torch.tensor([[0., 1.], [1., 0.]])
    with pytest.raises(RuntimeError, match="CUDA"):
        graph_operators([torch.zeros(2, 1), A], 1, True)
    with pytest.raises(RuntimeError, match="CUDA"):
        graph_oper()(torch.zeros(1, 2, 2, 3), torch.zeros(1, 1, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        GNN_simple(0, 2, 3, 5)([torch.zeros(1, 5, 2), torch.zeros(1, 2, 2, 3)], torch.tensor([2]), None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hgnn-2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_aliases_resolve_to_this_package():
    import hgnn_b200
    hgnn_b200.install_aliases(force=True)
    import functions.batching as fb
    import models.gnns.model_mnb as mm
    assert fb.__name__.startswith("hgnn_b200") and mm.__name__.startswith("hgnn_b200")
    import sys
    for k in [k for k in sys.modules if k.split(".")[0] in ("functions", "models")]:
        del sys.modules[k]
