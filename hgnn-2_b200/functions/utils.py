"""Mirror of the reference's functions/utils.py.

``graph_op`` / ``Pmul`` (:24-81) are the functional twins of the ``graph_oper`` / ``P_multi``
modules and run on the same CUDA kernels.  The meters and target normalisation (:84-146) are
host-side bookkeeping kept so that scripts/train_mnb.py / test_mnb.py run unchanged.
"""
import torch

from ..models.layers.layers_mnb import P_multi, graph_oper

_gop, _pmul = graph_oper(), P_multi()


def graph_op(A, X):
    """(bs,N,N,J) x (bs,F,N) -> (bs, J*F, N)   [reference functions/utils.py:24-52]."""
    return _gop(A, X)


def Pmul(P, X):
    """(bs,N,M) x (bs,F,M) -> (bs,F,N)   [reference functions/utils.py:55-81]."""
    return _pmul(P, X)


def data_stats(data):
    """reference :106-113 (std gets +1e-5)."""
    return (torch.min(data).item(), torch.max(data).item(), torch.mean(data).item(),
            1e-5 + torch.std(data).item())


def normalize_data(data, mean=None, std=None):
    """reference :84-95."""
    if mean is None or std is None:
        _, _, mean, std = data_stats(data)
    return data - mean if std < 1e-5 else (data - mean) / std


def evaluation(pred, target):
    """Mean absolute error (reference :98-103)."""
    return torch.mean(torch.abs(pred - target))


class AverageMeter():
    """reference :115-132."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class RunningAverage():
    """reference :134-146 - note the inverted momentum: val = 0.9*new + 0.1*old."""

    def __init__(self, momentum=0.1):
        self.momentum = momentum
        self.val = 0.0

    def update(self, val):
        self.val = val if self.val == 0.0 else (1 - self.momentum) * val + self.momentum * self.val
