"""hgnn-2_b200: B200-native (sm_100a) implementation of the HGNN-2 aggregation hot path.

Drop-in surface (same module paths / names / signatures as the reference):

    hgnn_b200.models.gnns.model_mnb.{GNN_simple, GNN_lg}
    hgnn_b200.models.layers.layers_mnb.{layer_simple, layer_last, layer_with_lg_1/2/3, layer_last_lg,
                                        graph_oper, P_multi}
    hgnn_b200.models.layers.batch_normalization.BN
    hgnn_b200.models.compnets.model_ccn.{CCN_1D, CCN_2D}
    hgnn_b200.functions.{operators, batching, utils, contraction, utils_ccn, data_generator, logs}

``install_aliases()`` registers these under the reference's top-level names ``models`` and
``functions`` so that the reference's unmodified scripts (scripts/train_mnb.py, train_ccn.py,
main_*.py) import this implementation instead.

Host bookkeeping lives in Python; every arithmetic op on features runs in the hand-written CUDA
kernels of ``csrc/`` behind the C ABI declared in ``include/hgnn_b200.h``.  No CPU fallback: the
import fails if ``libhgnn_b200.so`` is missing and every op raises without a CUDA device.
"""
import importlib
import sys

__version__ = "0.1.0"

from . import _lib  # noqa: F401,E402  (fails loudly when the CUDA library is missing)


def install_aliases(force=False):
    """Make ``import models.gnns.model_mnb`` / ``from functions import batching`` resolve here."""
    pkgs = ["functions", "functions.operators", "functions.batching", "functions.utils",
            "functions.logs", "functions.contraction", "functions.utils_ccn",
            "functions.data_generator", "models", "models.gnns", "models.gnns.model_mnb",
            "models.layers", "models.layers.layers_mnb", "models.layers.batch_normalization",
            "models.layers.gru_update", "models.compnets", "models.compnets.model_ccn"]
    for name in pkgs:
        if name in sys.modules and not force:
            owner = getattr(sys.modules[name], "__name__", "")
            if not owner.startswith(__name__):
                raise RuntimeError("module %r is already imported from elsewhere; call "
                                   "install_aliases(force=True) to override" % name)
        sys.modules[name] = importlib.import_module(__name__ + "." + name)
    return pkgs


def launch_count():
    """Number of kernels launched through the C ABI so far (bench.py's ``gpu_launches``)."""
    return _lib.total_launches()
