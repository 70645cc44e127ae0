"""``GRUUpdate`` / ``Identity`` (reference: models/layers/gru_update.py:17-42).

The reference constructs one of these in ``layer_simple.__init__`` (layers_mnb.py:38-41) but never
calls it (the call at :67 is commented out), so they are off the hot path.  Only what the drop-in
needs is kept: constructor arguments, the parameter names ``ih`` / ``hh`` (so ``state_dict`` keys and
whole-module pickles of ``GNN_simple(..., gru=True)`` stay compatible) and the gate arithmetic.
"""
import torch
import torch.nn as nn


class GRUUpdate(nn.Module):
    """Gated update o = (1 - z) * n + z * h with reset / update / candidate gates."""

    def __init__(self, fmap_in, fmap_out):
        super(GRUUpdate, self).__init__()
        self.fmap_out = fmap_out
        self.ih = nn.Linear(fmap_in, 3 * fmap_out)     # input  -> (reset, update, candidate)
        self.hh = nn.Linear(fmap_out, 3 * fmap_out)    # hidden -> (reset, update, candidate)

    def forward(self, i, h):
        F = self.fmap_out
        gi, gh = self.ih(i), self.hh(h)
        reset = torch.sigmoid(gi[..., :F] + gh[..., :F])
        update = torch.sigmoid(gi[..., F:2 * F] + gh[..., F:2 * F])
        cand = torch.tanh(torch.addcmul(gi[..., 2 * F:], reset, gh[..., 2 * F:]))
        return torch.lerp(cand, h, update)


class Identity(nn.Module):
    """Pass-through used when ``gru`` is False."""

    def forward(self, emb_in, emb_update):
        return emb_update
