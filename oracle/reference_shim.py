"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Makes the read-only reference checkout (``/root/reference``) importable in THIS container so that
``oracle/make_golden.py`` can generate golden vectors and ``tests/test_oracle_vs_reference.py`` can
pin the restatement in ``oracle/hgnn_oracle.py`` against the real code.  The reference does not
exist on the GPU box; everything that runs there uses the committed fixtures in ``tests/golden/``.

Three shims are needed on a modern stack:
  1. ``functions/logs.py:11-14`` imports matplotlib (absent here) -> stub modules.
  2. ``functions/utils_ccn.py:80-82`` tests an empty ``nonzero()`` with
     ``ind_j.shape == torch.Size([0])`` which is no longer how torch reports it ([0, 1]);
     the patched method tests ``numel() == 0`` instead.  Nothing else is touched.
  3. nothing on the hot path needs rdkit; ``preprocessing.preprocessing`` is not imported.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HGNN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "functions"))


class _Isolated:
    """Context manager: temporarily resolve ``functions.*`` / ``models.*`` to the reference."""

    _PREFIXES = ("functions", "models", "scripts", "preprocessing")

    def __enter__(self):
        self._saved = {k: v for k, v in sys.modules.items()
                       if k.split(".")[0] in self._PREFIXES}
        for k in self._saved:
            del sys.modules[k]
        sys.path.insert(0, REFERENCE_ROOT)
        return self

    def __exit__(self, *exc):
        sys.path.remove(REFERENCE_ROOT)
        mine = {k: v for k, v in sys.modules.items() if k.split(".")[0] in self._PREFIXES}
        for k in mine:
            del sys.modules[k]
        sys.modules.update(self._saved)
        self.modules = mine
        return False


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


_CACHE = {}


def load():
    """Import the reference's hot-path modules; returns a namespace of module objects."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _stub_matplotlib()
    import torch
    with _Isolated() as iso:
        import functions.operators as operators
        import functions.batching as batching
        import functions.utils as utils
        import functions.contraction as contraction
        import functions.utils_ccn as utils_ccn
        import models.layers.layers_mnb as layers_mnb
        import models.layers.batch_normalization as batch_normalization
        import models.gnns.model_mnb as model_mnb
        import models.compnets.model_ccn as model_ccn

    def _get_chi(self, i, j):  # functions/utils_ccn.py:66-91 with the empty-nonzero test fixed
        di = self.deg[i].item()
        dj = self.deg[j].item()
        chi = torch.zeros(di, dj)
        for k in range(di):
            ind_i = self.neighbors[i][k].item()
            ind_j = (self.neighbors[j] == ind_i).nonzero()
            if ind_j.numel() != 0:
                chi[k, ind_j.item()] = 1
        return chi

    utils_ccn.CompnetUtils._get_chi = _get_chi
    ns = types.SimpleNamespace(
        operators=operators, batching=batching, utils=utils, contraction=contraction,
        utils_ccn=utils_ccn, layers_mnb=layers_mnb, batch_normalization=batch_normalization,
        model_mnb=model_mnb, model_ccn=model_ccn)
    ns.sys_modules = iso.modules       # {name: module} of the reference, for pickling by class path
    _CACHE["ns"] = ns
    return ns
