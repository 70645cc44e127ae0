"""Padding-aware batch-norm with scalar affine - the layer epilogue.

Mirror of the reference's models/layers/batch_normalization.py (``BN`` :23-43,
``sb_normalization`` :65-77).  Same parameters (0-dim ``weight`` / ``bias`` ~ N(0, 0.1)), same
running-statistics rule (``0.9*batch + 0.1*running``, :37-38), same train/eval switch - computed by
the CUDA kernels of csrc/bn.cu on packed rows.  ``running_mean`` / ``running_std`` are
non-persistent buffers: they follow ``.cuda()`` and whole-module pickles (functions/logs.py:99-111)
but stay out of ``state_dict`` exactly like the reference's plain attributes.
"""
import torch
import torch.nn as nn

from ... import ops
from ..._lib import require_cuda


def _offsets(N_batch, device):
    n = N_batch.to(device=device, dtype=torch.int32)
    off = torch.zeros(n.numel() + 1, dtype=torch.int32, device=device)
    off[1:] = torch.cumsum(n, 0)
    return off


class BN(nn.Module):
    def __init__(self, n_features, scale=0.1):
        super(BN, self).__init__()
        self.n_features = n_features
        self.weight = nn.Parameter(torch.zeros(()).normal_(0, scale))
        self.bias = nn.Parameter(torch.zeros(()).normal_(0, scale))
        self.register_buffer("running_mean", torch.zeros(n_features), persistent=False)
        self.register_buffer("running_std", torch.zeros(n_features), persistent=False)
        self.momentum = 0.1

    def __setstate__(self, state):
        """Whole-module pickles (functions/logs.py:99-123).  A module pickled by the REFERENCE keeps its
        running statistics as plain attributes (batch_normalization.py:30-31) and has no ``n_features``:
        re-home them as the non-persistent buffers this class uses, so they follow ``.cuda()``."""
        super(BN, self).__setstate__(state)
        for k in ("running_mean", "running_std"):
            if k in self.__dict__:
                v = self.__dict__.pop(k)
                self.register_buffer(k, v.detach().float().contiguous(), persistent=False)
        if "n_features" not in self.__dict__:
            self.n_features = int(self.running_mean.numel())

    def forward_rows(self, Z):
        """Packed rows (R, F) -> (normalised rows, stats)."""
        return ops.BatchNormRows.apply(Z, self.weight, self.bias, self, self.training)

    def forward(self, X, N_batch, mask=None):
        """Reference signature: X (bs, F, Nmax) padded; padded slots come out as
        ``weight*(0-mean)/std + bias`` exactly like batch_normalization.py:75,43."""
        require_cuda()
        bs, F, Nmax = X.shape
        off = _offsets(N_batch, X.device)
        R = int(N_batch.sum().item())
        Zp = ops.PackRows.apply(X, off, R)
        Y, stats = self.forward_rows(Zp)
        return ops.UnpackRows.apply(Y, off, bs, Nmax, stats[3 * F:].contiguous())


def sb_normalization(H, N_batch, mask=None, mean=None, std=None):
    """batch_normalization.py:65-77 - normalisation without the affine part.
    Returns (H_normalised, mean, std)."""
    require_cuda()
    bs, F, Nmax = H.shape
    bn = BN(F).to(H.device)
    with torch.no_grad():
        bn.weight.fill_(1.0)
        bn.bias.fill_(0.0)
    if torch.is_tensor(mean) and torch.is_tensor(std):
        bn.running_mean, bn.running_std = mean.float().contiguous(), std.float().contiguous()
        bn.eval()
    off = _offsets(N_batch, H.device)
    Zp = ops.PackRows.apply(H, off, int(N_batch.sum().item()))
    Y, stats = ops.BatchNormRows.apply(Zp, bn.weight.detach(), bn.bias.detach(), bn, bn.training)
    out = ops.UnpackRows.apply(Y, off, bs, Nmax, stats[3 * F:].contiguous())
    return out, stats[:F], stats[F:2 * F]
