"""GNN / LGNN layers on the fused CUDA path.

Mirror of the reference's models/layers/layers_mnb.py: same class names, constructor arguments,
parameter names/shapes (``cv1..cv4`` Conv1d(k=1) weights ``(Fout, Cin, 1)``, scalar ``bn1/bn2``,
``fc``), same ``forward`` signatures and state tuples.  The reference computes every layer as
``bs*(K)`` dense ``torch.mm`` + ``torch.cat`` + two Conv1d + ReLU + BN (layers_mnb.py:52-69,
189-225, 256-290, 322-358); here each *side* of a layer (node update, edge update, readout) is one
fused kernel (csrc/side.cu) over the block-diagonal CSR pack.

Every layer also has a ``forward_packed`` used by the models so that a stack of layers converts the
reference's padded ``(bs, F, Nmax)`` layout to packed rows once, not per layer.
"""
import torch
import torch.nn as nn

from ... import ops
from ..._lib import require_cuda
from ...pack import BatchPack, OperatorHandle, dense_to_csr, resolve_pack
from .batch_normalization import BN
from .gru_update import GRUUpdate, Identity


def _side_cfgs(pack):
    """(node side, edge side) static descriptors, cached on the pack."""
    cfg = getattr(pack, "_side_cfgs", None)
    if cfg is None:
        if pack.dual:
            node = ops.SideCfg(pack.Rn, pack.node_ops(), pack.node_ops_T(), pack.p, pack.pt, pack.Rm)
            edge = ops.SideCfg(pack.Rm, pack.edge_ops(), pack.edge_ops_T(), pack.pt, pack.p, pack.Rn)
        else:
            node = ops.SideCfg(pack.Rn, pack.node_ops(), pack.node_ops_T())
            edge = None
        cfg = pack._side_cfgs = (node, edge)
    return cfg


def _init_convs(convs, scale=0.1):
    for l in convs:
        l.weight.data.normal_(0, scale)
        l.bias.data.normal_(0, scale)


def _pack_nodes(pack, X):
    return ops.PackRows.apply(X, pack.node_off, pack.Rn)


def _pack_edges(pack, XL):
    return ops.PackRows.apply(XL, pack.edge_off, pack.Rm)


def _unpack_nodes(pack, Y, stats=None):
    F = Y.shape[1]
    fill = None if stats is None else stats[3 * F:].contiguous()
    return ops.UnpackRows.apply(Y, pack.node_off, pack.bs, pack.Nmax, fill)


def _unpack_edges(pack, Y, stats=None):
    F = Y.shape[1]
    fill = None if stats is None else stats[3 * F:].contiguous()
    return ops.UnpackRows.apply(Y, pack.edge_off, pack.bs, pack.Emax, fill)


class layer_simple(nn.Module):
    """Layer of the power GNN (reference layers_mnb.py:25-69): both conv branches are ReLU'd."""

    def __init__(self, feature_maps, J, gru):
        super(layer_simple, self).__init__()
        self.n_inputs = feature_maps[0]
        self.n_outputs = feature_maps[1]
        self.gop = graph_oper()
        self.cv1 = torch.nn.Conv1d(J * self.n_inputs, self.n_outputs, 1)
        self.cv2 = torch.nn.Conv1d(J * self.n_inputs, self.n_outputs, 1)
        self.update = GRUUpdate(self.n_inputs, 2 * self.n_outputs) if gru == True else Identity()  # noqa: E712
        self.bn1 = BN(2 * self.n_outputs)
        _init_convs([self.cv1, self.cv2])

    def forward_packed(self, Xp, pack):
        node, _ = _side_cfgs(pack)
        return ops.side_update(Xp, None, self.cv2.weight, self.cv2.bias, self.cv1.weight,
                                    self.cv1.bias, self.bn1.weight, self.bn1.bias, node, 0, self.bn1,
                                    self.training)

    def forward(self, state, N_batch, mask):
        require_cuda()
        X, W = state
        pack = resolve_pack(W, N_batch=N_batch)
        Y, stats = self.forward_packed(_pack_nodes(pack, X), pack)
        return (_unpack_nodes(pack, Y, stats), W)


class layer_last(nn.Module):
    """Readout of the power GNN (reference layers_mnb.py:72-95)."""

    def __init__(self, feature_maps, J):
        super(layer_last, self).__init__()
        self.n_inputs = feature_maps[0]
        self.n_outputs = feature_maps[1]
        self.gop = graph_oper()
        self.fc = torch.nn.Conv1d(J * self.n_inputs, self.n_outputs, 1)
        _init_convs([self.fc])

    def forward_packed(self, Xp, pack):
        node, _ = _side_cfgs(pack)
        return ops.Readout.apply(Xp, None, self.fc.weight, self.fc.bias, node, pack.node_off, pack.pad_n,
                                 pack.bs)

    def forward(self, state, N_batch, mask):
        require_cuda()
        X, W = state
        pack = resolve_pack(W, N_batch=N_batch)
        return self.forward_packed(_pack_nodes(pack, X), pack)


class _layer_with_lg(nn.Module):
    """Shared body of layer_with_lg_{1,2,3} (reference layers_mnb.py:157-358).  ``order`` decides
    which state feeds which update and therefore the conv widths (:172-177, :239-244, :305-310)."""
    order = 0

    def __init__(self, feature_maps, J):
        super(_layer_with_lg, self).__init__()
        self.n_inputs = feature_maps[0]
        self.n_edges = feature_maps[1]
        self.n_outputs = feature_maps[2]
        self.gop = graph_oper()
        self.pmul = P_multi()
        fn, fe, h = self.n_inputs, self.n_edges, self.n_outputs
        node_in = J * fn + (4 * h if self.order == 2 else 2 * fe)
        edge_in = J * fe + (4 * h if self.order == 1 else 2 * fn)
        self.cv1 = torch.nn.Conv1d(node_in, h, 1)
        self.cv2 = torch.nn.Conv1d(node_in, h, 1)
        self.bn1 = BN(2 * h)
        self.cv3 = torch.nn.Conv1d(edge_in, h, 1)
        self.cv4 = torch.nn.Conv1d(edge_in, h, 1)
        self.bn2 = BN(2 * h)
        _init_convs([self.cv1, self.cv2, self.cv3, self.cv4])

    def _node(self, Xp, edge_state, cfg):
        # cat(cv2 branch [no ReLU], relu(cv1 branch)) -> bn1   (layers_mnb.py:206-212)
        return ops.side_update(Xp, edge_state, self.cv2.weight, self.cv2.bias, self.cv1.weight,
                                    self.cv1.bias, self.bn1.weight, self.bn1.bias, cfg,
                                    self.n_outputs, self.bn1, self.training)

    def _edge(self, XLp, node_state, cfg):
        # cat(cv4 branch [no ReLU], relu(cv3 branch)) -> bn2   (layers_mnb.py:217-223)
        return ops.side_update(XLp, node_state, self.cv4.weight, self.cv4.bias, self.cv3.weight,
                                    self.cv3.bias, self.bn2.weight, self.bn2.bias, cfg,
                                    self.n_outputs, self.bn2, self.training)

    def forward_packed(self, Xp, XLp, pack):
        node, edge = _side_cfgs(pack)
        if self.order == 1:        # edges see the NEW node state (:214-215)
            zbn1, s1 = self._node(Xp, XLp, node)
            zdbn1, s2 = self._edge(XLp, zbn1, edge)
        elif self.order == 2:      # nodes see the NEW edge state (:277-278)
            zdbn1, s2 = self._edge(XLp, Xp, edge)
            zbn1, s1 = self._node(Xp, zdbn1, node)
        else:                      # both from the old states (:333-340)
            zbn1, s1 = self._node(Xp, XLp, node)
            zdbn1, s2 = self._edge(XLp, Xp, edge)
        return zbn1, zdbn1, s1, s2

    def forward(self, state, N_batch, mask, E_batch, mask_lg):
        require_cuda()
        X, XL, W, WL, Pm, Pd = state
        pack = resolve_pack(W, WL, Pm, Pd, N_batch, E_batch)
        zbn1, zdbn1, s1, s2 = self.forward_packed(_pack_nodes(pack, X), _pack_edges(pack, XL), pack)
        return (_unpack_nodes(pack, zbn1, s1), _unpack_edges(pack, zdbn1, s2), W, WL, Pm, Pd)


class layer_with_lg_1(_layer_with_lg):
    order = 1


class layer_with_lg_2(_layer_with_lg):
    order = 2


class layer_with_lg_3(_layer_with_lg):
    order = 3


class layer_last_lg(nn.Module):
    """Readout of the LGNN (reference layers_mnb.py:361-388)."""

    def __init__(self, feature_maps, J):
        super(layer_last_lg, self).__init__()
        self.n_inputs = feature_maps[0]
        self.n_outputs = feature_maps[1]
        self.gop = graph_oper()
        self.pmul = P_multi()
        self.fc = torch.nn.Conv1d((J + 2) * self.n_inputs, self.n_outputs, 1)
        _init_convs([self.fc])

    def forward_packed(self, Xp, XLp, pack):
        node, _ = _side_cfgs(pack)
        return ops.Readout.apply(Xp, XLp, self.fc.weight, self.fc.bias, node, pack.node_off, pack.pad_n,
                                 pack.bs)

    def forward(self, state, N_batch, mask):
        require_cuda()
        X, XL, W, WL, Pm, Pd = state
        E_batch = None
        if torch.is_tensor(WL):   # dense compatibility path: every padded edge slot counts as a row
            E_batch = torch.full((WL.shape[0],), WL.shape[1], dtype=torch.int64)
        pack = resolve_pack(W, WL, Pm, Pd, N_batch, E_batch)
        return self.forward_packed(_pack_nodes(pack, X), _pack_edges(pack, XL), pack)


# --------------------------------------------------------------------------------------------
# stand-alone operators (the reference's graph_oper / P_multi modules, layers_mnb.py:391-434)
# --------------------------------------------------------------------------------------------


def _full_offsets(bs, n, device):
    return torch.arange(bs + 1, dtype=torch.int32, device=device) * n


class graph_oper(nn.Module):
    """"gmul": out[b, j*F+f, v] = sum_u A[b,v,u,j] X[b,f,u] (reference layers_mnb.py:391-411)."""

    def forward(self, A, X):
        require_cuda()
        if isinstance(A, OperatorHandle):
            pack = A.pack
            if A.name == "W":
                off, R, Nmax, descs, descs_T = pack.node_off, pack.Rn, pack.Nmax, pack.node_ops(), pack.node_ops_T()
            elif A.name == "WL":
                off, R, Nmax, descs, descs_T = pack.edge_off, pack.Rm, pack.Emax, pack.edge_ops(), pack.edge_ops_T()
            else:
                raise RuntimeError("graph_oper expects a W or WL handle, got %s" % A.name)
            bs = pack.bs
        else:
            if not A.is_cuda:
                raise RuntimeError("hgnn_b200: graph_oper needs CUDA tensors (no CPU fallback)")
            bs, Nmax, _, K = A.shape
            off, R = _full_offsets(bs, Nmax, A.device), bs * Nmax
            Ad = A.detach()
            descs = [dense_to_csr(Ad[:, :, :, k], None, bs, off, off, R).desc() for k in range(K)]
            descs_T = [dense_to_csr(Ad[:, :, :, k].transpose(1, 2), None, bs, off, off, R).desc()
                       for k in range(K)]
        Xp = ops.PackRows.apply(X, off, R)
        Y = ops.Gmul.apply(Xp, descs, descs_T, R)
        return ops.UnpackRows.apply(Y, off, bs, Nmax, None)


class P_multi(nn.Module):
    """out[b,f,v] = sum_e P[b,v,e] X[b,f,e] (reference layers_mnb.py:414-434)."""

    def forward(self, P, X):
        require_cuda()
        if isinstance(P, OperatorHandle):
            pack = P.pack
            second = P.name == "Pd"
            if P.name not in ("Pm", "Pd"):
                raise RuntimeError("P_multi expects a Pm or Pd handle, got %s" % P.name)
            if P.transposed:   # (bs, M, N): node features -> line-graph nodes
                off_out, R_out, n_out = pack.edge_off, pack.Rm, pack.Emax
                off_in, R_in = pack.node_off, pack.Rn
                fwd, bwd = pack.pt.desc(second), pack.p.desc(second)
            else:
                off_out, R_out, n_out = pack.node_off, pack.Rn, pack.Nmax
                off_in, R_in = pack.edge_off, pack.Rm
                fwd, bwd = pack.p.desc(second), pack.pt.desc(second)
            bs = pack.bs
        else:
            if not P.is_cuda:
                raise RuntimeError("hgnn_b200: P_multi needs CUDA tensors (no CPU fallback)")
            bs, n_out, n_in = P.shape
            off_out, R_out = _full_offsets(bs, n_out, P.device), bs * n_out
            off_in, R_in = _full_offsets(bs, n_in, P.device), bs * n_in
            Pd_ = P.detach()
            fwd = dense_to_csr(Pd_, None, bs, off_out, off_in, R_out).desc()
            bwd = dense_to_csr(Pd_.transpose(1, 2), None, bs, off_in, off_out, R_in).desc()
        Xp = ops.PackRows.apply(X, off_in, R_in)
        Y = ops.Gmul.apply(Xp, [fwd], [bwd], R_out)
        return ops.UnpackRows.apply(Y, off_out, bs, n_out, None)
