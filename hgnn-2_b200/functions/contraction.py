"""The 18 second-order contractions - mirror of the reference's functions/contraction.py.

``collapse6to3(F)`` (:106-121) keeps its signature: F is ``(C, n, n, n, n, n)``, the result is
``(n, n, 18*C)`` with contraction k, channel c at column ``k*C + c``.  The reference evaluates it
with 18 permutes, diagonal masks and triple sums over n^5*C elements; here one CUDA kernel
(csrc/ccn.cu) evaluates all 18 and a second one the backward.  The hot path
(``CompnetUtils.update_F``) does not build the rank-6 tensor at all - see utils_ccn.py.
"""
import torch

from .. import ops
from .._lib import require_cuda

NUM_CONTRACTIONS = 18


def collapse6to3(F):
    require_cuda()
    assert all(F.shape[1] == F.shape[i] for i in range(1, F.dim()))
    if not F.is_cuda:
        raise RuntimeError("hgnn_b200: collapse6to3 needs a CUDA tensor (no CPU fallback)")
    out = ops.Collapse6to3.apply(F)
    assert out.shape == (F.shape[1], F.shape[1], F.shape[0] * NUM_CONTRACTIONS)
    return out


def _split(F_j, lo, hi, F):
    C = F.shape[0]
    return [F_j[:, :, k * C:(k + 1) * C] for k in range(lo, hi)]


def _c6to2_111(F):
    """Cases 1-5 (reference :44-61).  F: (n,n,n,n,n,C) as in the reference."""
    return _split(collapse6to3(F.permute(5, 0, 1, 2, 3, 4)), 0, 5, F.permute(5, 0, 1, 2, 3, 4))


def _c6to2_12(F):
    """Cases 6-15 (reference :64-85)."""
    return _split(collapse6to3(F.permute(5, 0, 1, 2, 3, 4)), 5, 15, F.permute(5, 0, 1, 2, 3, 4))


def _c6to2_3(F):
    """Cases 16-18 (reference :88-103)."""
    return _split(collapse6to3(F.permute(5, 0, 1, 2, 3, 4)), 15, 18, F.permute(5, 0, 1, 2, 3, 4))
