"""``GRUUpdate`` / ``Identity`` (reference: models/layers/gru_update.py:17-42).

The reference constructs these in ``layer_simple.__init__`` (layers_mnb.py:38-41) but never calls
them (the call at :67 is commented out), so they are off the hot path: only the constructor
arguments and parameter shapes are kept, so ``GNN_simple(..., gru=True)`` still builds and pickles.
"""
import torch
import torch.nn as nn


class GRUUpdate(nn.Module):
    def __init__(self, fmap_in, fmap_out):
        super(GRUUpdate, self).__init__()
        self.ih = nn.Linear(fmap_in, 3 * fmap_out)
        self.hh = nn.Linear(fmap_out, 3 * fmap_out)

    def forward(self, i, h):
        gi, gh = self.ih(i), self.hh(h)
        r_i, z_i, n_i = gi.chunk(3, -1)
        r_h, z_h, n_h = gh.chunk(3, -1)
        z = torch.sigmoid(z_i + z_h)
        n = torch.tanh(n_i + torch.sigmoid(r_i + r_h) * n_h)
        return (1 - z) * n + z * h


class Identity(nn.Module):
    def forward(self, emb_in, emb_update):
        return emb_update
