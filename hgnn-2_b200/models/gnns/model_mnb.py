"""``GNN_simple`` / ``GNN_lg`` - mirror of the reference's models/gnns/model_mnb.py (:19-66, :69-129).

Same constructor signatures, module names (``layer0``, ``layer{i}``, ``layerlast``), attributes
(``.dual``, ``.J`` read by scripts/train_mnb.py:29-30) and forward signatures.  The forward converts
the padded ``(bs, F, Nmax)`` inputs to packed rows once, runs the layer stack on the fused kernels
and returns ``(bs, dim_output)`` like the reference.
"""
import torch.nn as nn

from ... import engine
from ..._lib import require_cuda
from ...pack import PackTensor, resolve_pack
from ..layers import layers_mnb


def _needs_layer_autograd(model, X, XL):
    """The model-level engine differentiates w.r.t. the parameters and X in train mode only.  The two cases it does
    not cover take the per-layer autograd path, like the reference's plain autograd would: a gradient asked for the
    edge features XL, and eval mode with gradients enabled (input-gradient analysis of a trained model)."""
    import torch
    if XL is not None and torch.is_tensor(XL) and XL.requires_grad and torch.is_grad_enabled():
        return True
    if not model.training and torch.is_grad_enabled():
        if (torch.is_tensor(X) and X.requires_grad) or (XL is not None and torch.is_tensor(XL) and XL.requires_grad):
            return True
    return False


class GNN_simple(nn.Module):
    """Power GNN.  ``task`` and ``gru`` are accepted and unused, as in the reference (:42)."""

    def __init__(self, task, n_features, n_layers, dim_input, dim_output=1, J=1, gru=False):
        super(GNN_simple, self).__init__()
        self.dual = False
        self.J = J
        self.gru = False
        self.n_features = n_features
        self.n_layers = n_layers
        self.n_outputs = dim_output
        self.featuremap_in = [dim_input, n_features]
        self.featuremap_mi = [2 * n_features, n_features]
        self.featuremap_end = [2 * n_features, dim_output]
        self.layer0 = layers_mnb.layer_simple(self.featuremap_in, J + 2, gru)
        for i in range(n_layers - 2):
            self.add_module('layer{}'.format(i + 1), layers_mnb.layer_simple(self.featuremap_mi, J + 2, gru))
        self.layerlast = layers_mnb.layer_last(self.featuremap_end, J + 2)

    def forward(self, state, N_batch, mask):
        require_cuda()
        X, W = state
        pack = resolve_pack(W, N_batch=N_batch)
        if engine.supported(self) and not _needs_layer_autograd(self, X, None):
            return engine.run_model(self, pack, layers_mnb._pack_nodes(pack, X), None)   # model-level engine (csrc/engine.cu)
        cur, _ = self.layer0.forward_packed(layers_mnb._pack_nodes(pack, X), pack)
        for i in range(self.n_layers - 2):
            cur, _ = self._modules['layer{}'.format(i + 1)].forward_packed(cur, pack)
        return self.layerlast.forward_packed(cur, pack)


class GNN_lg(nn.Module):
    """GNN on the line graph with the non-backtracking operator; ``order`` in {1, 2, 3} selects
    layer_with_lg_{1,2,3} (reference :102-120; any other value falls through to 3 there too)."""

    def __init__(self, task, n_features, n_layers, dim_input, dim_output=1, J=1, order=1):
        super(GNN_lg, self).__init__()
        self.dual = True
        self.J = J
        self.n_features = n_features
        self.n_layers = n_layers
        self.n_outputs = dim_output
        self.order = order
        self.featuremap_in = [dim_input, 1, n_features]
        self.featuremap_mi = [2 * n_features, 2 * n_features, n_features]
        self.featuremap_end = [2 * n_features, dim_output]
        cls = {1: layers_mnb.layer_with_lg_1, 2: layers_mnb.layer_with_lg_2}.get(order, layers_mnb.layer_with_lg_3)
        self.layer0 = cls(self.featuremap_in, J + 2)
        for i in range(n_layers - 2):
            self.add_module('layer{}'.format(i + 1), cls(self.featuremap_mi, J + 2))
        self.layerlast = layers_mnb.layer_last_lg(self.featuremap_end, J + 2)

    def forward(self, state, N_batch, mask, E_batch, mask_lg):
        require_cuda()
        X, XL, W, WL, Pm, Pd = state
        pack = resolve_pack(W, WL, Pm, Pd, N_batch, E_batch)
        Xp = layers_mnb._pack_nodes(pack, X)
        # XL straight from prepare_batch IS the pack's line-graph degree (functions/batching.py:171): use the
        # device copy the pack already holds, and let the engine collapse the identical phantom rows
        degree = PackTensor.pack_of(XL) is pack and not pack.generic and not XL.requires_grad
        XLp = pack.dl.view(-1, 1) if degree else layers_mnb._pack_edges(pack, XL)
        if engine.supported(self) and not _needs_layer_autograd(self, X, XL):
            return engine.run_model(self, pack, Xp, XLp, xl_is_degree=degree)   # model-level engine (csrc/engine.cu)
        Xp, XLp, _, _ = self.layer0.forward_packed(Xp, XLp, pack)
        for i in range(self.n_layers - 2):
            Xp, XLp, _, _ = self._modules['layer{}'.format(i + 1)].forward_packed(Xp, XLp, pack)
        return self.layerlast.forward_packed(Xp, XLp, pack)
