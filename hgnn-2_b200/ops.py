"""torch.autograd.Function wrappers over the C ABI (include/hgnn_b200.h).

PyTorch is plumbing here: it owns device memory, the stream and the autograd tape; every arithmetic
step on features is one of the hand-written kernels.  All feature tensors are "packed rows"
(R, F) fp32 (data layout).
"""
import ctypes

import torch

from . import _lib
from ._lib import SideT, call, fptr, iptr, make_ops, stream, workspace


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------------------------
# layout
# --------------------------------------------------------------------------------------------


class PackRows(torch.autograd.Function):
    """dense (bs, F, Nmax) -> packed (R, F)   [boundary of functions/batching.py:77-185]."""

    @staticmethod
    def forward(ctx, dense, off, R):
        dense = _f32c(dense)
        bs, F, Nmax = dense.shape
        out = torch.empty(R, F, device=dense.device)
        call("hgnn_pack_rows", fptr(dense), bs, F, Nmax, iptr(off), fptr(out), stream())
        ctx.off, ctx.dims = off, (bs, F, Nmax)
        return out

    @staticmethod
    def backward(ctx, g):
        bs, F, Nmax = ctx.dims
        g = _f32c(g)
        out = torch.empty(bs, F, Nmax, device=g.device)
        call("hgnn_unpack_rows", fptr(g), bs, F, Nmax, iptr(ctx.off), None, fptr(out), stream())
        return out, None, None


class UnpackRows(torch.autograd.Function):
    """packed (R, F) -> dense (bs, F, Nmax); padded slots get ``pad_fill[f]`` (the value the
    reference's BN leaves there, batch_normalization.py:75) or 0."""

    @staticmethod
    def forward(ctx, packed, off, bs, Nmax, pad_fill):
        packed = _f32c(packed)
        R, F = packed.shape
        out = torch.empty(bs, F, Nmax, device=packed.device)
        call("hgnn_unpack_rows", fptr(packed), bs, F, Nmax, iptr(off),
             None if pad_fill is None else fptr(pad_fill), fptr(out), stream())
        ctx.off, ctx.dims = off, (bs, F, Nmax, R)
        return out

    @staticmethod
    def backward(ctx, g):
        bs, F, Nmax, R = ctx.dims
        g = _f32c(g)
        out = torch.empty(R, F, device=g.device)
        call("hgnn_pack_rows", fptr(g), bs, F, Nmax, iptr(ctx.off), fptr(out), stream())
        gfill = None
        if ctx.needs_input_grad[4]:
            # what reached the padded slots (only when a caller differentiates through them; the
            # models never do): everything minus the real rows
            gfill = g.sum(dim=(0, 2)) - out.sum(dim=0)
        return out, None, None, None, gfill


# --------------------------------------------------------------------------------------------
# stand-alone gmul  (graph_oper / P_multi)
# --------------------------------------------------------------------------------------------


class Gmul(torch.autograd.Function):
    """Y = [op_0 X | op_1 X | ...]  (layers_mnb.py:395-411, 418-434)."""

    @staticmethod
    def forward(ctx, X, descs, descs_T, R_out):
        X = _f32c(X)
        F = X.shape[1]
        ops, n = make_ops(descs)
        Y = torch.empty(R_out, n * F, device=X.device)
        call("hgnn_gmul_fwd", ops, n, R_out, F, fptr(X), fptr(Y), stream())
        ctx.descs_T, ctx.R_in, ctx.F = descs_T, X.shape[0], F
        return Y

    @staticmethod
    def backward(ctx, G):
        G = _f32c(G)
        ops, n = make_ops(ctx.descs_T)
        gX = torch.empty(ctx.R_in, ctx.F, device=G.device)
        call("hgnn_gmul_bwd", ops, n, ctx.R_in, ctx.F, fptr(G), fptr(gX), stream())
        return gX, None, None, None


# --------------------------------------------------------------------------------------------
# batch-norm on packed rows
# --------------------------------------------------------------------------------------------


def bn_stats_eval(F, weight, bias, running_mean, running_std):
    stats = torch.empty(4 * F, device=weight.device)
    call("hgnn_bn_stats_eval", F, fptr(weight), fptr(bias), fptr(running_mean), fptr(running_std),
         fptr(stats), stream())
    return stats


class BatchNormRows(torch.autograd.Function):
    """BN.forward on packed rows (batch_normalization.py:34-43).  Returns (Y, stats)."""

    @staticmethod
    def forward(ctx, Z, weight, bias, bn, training):
        Z = _f32c(Z)
        R, F = Z.shape
        if training:
            stats = torch.empty(4 * F, device=Z.device)
            ws, wsb = workspace(2 * F, Z.device)
            call("hgnn_bn_stats", fptr(Z), R, F, fptr(weight), fptr(bias), fptr(bn.running_mean),
                 fptr(bn.running_std), float(bn.momentum), fptr(stats), ws, wsb, stream())
        else:
            stats = bn_stats_eval(F, weight, bias, bn.running_mean, bn.running_std)
        Y = torch.empty_like(Z)
        call("hgnn_bn_apply", fptr(Z), R, F, fptr(stats), fptr(Y), stream())
        ctx.save_for_backward(Z, stats, weight)
        ctx.training = training
        ctx.set_materialize_grads(False)
        return Y, stats

    @staticmethod
    def backward(ctx, gY, gstats):
        Z, stats, weight = ctx.saved_tensors
        gY = torch.zeros_like(Z) if gY is None else _f32c(gY)
        R, F = Z.shape
        coef = torch.empty(3 * F + 2, device=Z.device)
        ws, wsb = workspace(2 * F, Z.device)
        gshift = None if gstats is None else _f32c(gstats[3 * F:])
        call("hgnn_bn_bwd_reduce", fptr(gY), fptr(Z), R, F, fptr(stats), fptr(weight),
             1 if ctx.training else 0, fptr(gshift), fptr(coef), ws, wsb, stream())
        gZ = torch.empty_like(Z)
        call("hgnn_side_bwd_pre", fptr(gY), fptr(Z), R, F, fptr(coef), F, fptr(gZ), None, ws, wsb, stream())
        return gZ, coef[3 * F].reshape(weight.shape), coef[3 * F + 1].reshape(weight.shape), None, None


# --------------------------------------------------------------------------------------------
# fused layer side
# --------------------------------------------------------------------------------------------


class SideCfg(object):
    """Static description of one layer side: which operators, which incidence orientation."""
    __slots__ = ("R", "ops", "ops_T", "p", "pt", "R_cross")

    def __init__(self, R, ops, ops_T, p=None, pt=None, R_cross=0):
        self.R, self.ops, self.ops_T, self.p, self.pt, self.R_cross = R, ops, ops_T, p, pt, R_cross


def _side_struct(cfg, Xs, Xc):
    ops, n = make_ops(cfg.ops)
    s = SideT()
    s.R, s.ops, s.n_ops = cfg.R, ops, n
    s.Xs, s.Fs = fptr(Xs), Xs.shape[1]
    if cfg.p is not None and Xc is not None:
        s.p_rowptr, s.p_col = iptr(cfg.p.rowptr), iptr(cfg.p.col)
        s.p_pm, s.p_pd = fptr(cfg.p.val), fptr(cfg.p.val2)
        s.Xc, s.Fc = fptr(Xc), Xc.shape[1]
    else:
        s.p_rowptr = s.p_col = s.p_pm = s.p_pd = s.Xc = None
        s.Fc = 0
    return s, ops   # keep `ops` alive while the struct is in use


def _side_backward(cfg, gPre, Xs, Xc, Wa, Wb, need_gxs, need_gxc):
    """Transposed gathers: returns (gXs, gXc, dWa, dWb)."""
    dev = gPre.device
    Fg = gPre.shape[1]
    Ha, Hb = (Wa.shape[0] if Wa is not None else 0), (Wb.shape[0] if Wb is not None else 0)
    Cin = (Wa if Wa is not None else Wb).shape[1]
    dWa = torch.empty(Ha, Cin, device=dev) if Ha else None
    dWb = torch.empty(Hb, Cin, device=dev) if Hb else None
    Fs = Xs.shape[1]
    opsT, n = make_ops(cfg.ops_T)
    gXs = torch.empty_like(Xs) if need_gxs else None
    ws, wsb = workspace(n * Fg * Fs, dev)
    call("hgnn_side_bwd_gather", opsT, n, cfg.R, fptr(gPre), Fg, fptr(Xs), Fs, fptr(Wa), Ha, fptr(Wb), Hb,
         Cin, 0, fptr(gXs), 0, fptr(dWa), fptr(dWb), ws, wsb, stream())
    gXc = None
    if Xc is not None:
        Fc = Xc.shape[1]
        pt = cfg.pt
        cross, nc = make_ops([pt.desc(False), pt.desc(True)])
        gXc = torch.empty_like(Xc) if need_gxc else None
        ws, wsb = workspace(2 * Fg * Fc, dev)
        call("hgnn_side_bwd_gather", cross, nc, cfg.R_cross, fptr(gPre), Fg, fptr(Xc), Fc, fptr(Wa), Ha,
             fptr(Wb), Hb, Cin, n * Fs, fptr(gXc), 0, fptr(dWa), fptr(dWb), ws, wsb, stream())
    return gXs, gXc, dWa, dWb


class SideUpdate(torch.autograd.Function):
    """One side of a GNN / LGNN layer, fused:  gather -> cat -> (cv_a | cv_b) -> ReLU -> BN.

    layers_mnb.py:58-68 (layer_simple: relu_from = 0), :200-212 / :214-223 (layer_with_lg_1 node /
    edge side: relu_from = h) and the order-2/3 variants.  Output = cat(cv_a branch, cv_b branch)
    after batch-norm; also returns the BN ``stats`` vector (mean, std, scale, shift)."""

    @staticmethod
    def forward(ctx, Xs, Xc, Wa, ba, Wb, bb, bn_w, bn_b, cfg, relu_from, bn, training):
        Xs = _f32c(Xs)
        Xc = _f32c(Xc) if Xc is not None else None
        Wa2, Wb2 = _f32c(Wa).view(Wa.shape[0], -1), _f32c(Wb).view(Wb.shape[0], -1)
        Ha, Hb = Wa2.shape[0], Wb2.shape[0]
        F = Ha + Hb
        dev = Xs.device
        side, keep = _side_struct(cfg, Xs, Xc)
        Z = torch.empty(cfg.R, F, device=dev)
        ws, wsb = workspace(2 * F, dev)
        if training:
            stats = torch.empty(4 * F, device=dev)
            call("hgnn_side_fwd", ctypes.byref(side), fptr(Wa2), fptr(ba), Ha, fptr(Wb2), fptr(bb), Hb,
                 relu_from, fptr(Z), fptr(bn_w), fptr(bn_b), fptr(bn.running_mean), fptr(bn.running_std),
                 float(bn.momentum), fptr(stats), ws, wsb, stream())
        else:
            call("hgnn_side_fwd", ctypes.byref(side), fptr(Wa2), fptr(ba), Ha, fptr(Wb2), fptr(bb), Hb,
                 relu_from, fptr(Z), None, None, None, None, 0.0, None, None, 0, stream())
            stats = bn_stats_eval(F, bn_w, bn_b, bn.running_mean, bn.running_std)
        Y = torch.empty_like(Z)
        call("hgnn_bn_apply", fptr(Z), cfg.R, F, fptr(stats), fptr(Y), stream())
        ctx.save_for_backward(Xs, Xc, Wa2, Wb2, Z, stats, bn_w)
        ctx.cfg, ctx.relu_from, ctx.training = cfg, relu_from, training
        ctx.wshapes = (Wa.shape, Wb.shape)
        ctx.set_materialize_grads(False)
        return Y, stats

    @staticmethod
    def backward(ctx, gY, gstats):
        Xs, Xc, Wa2, Wb2, Z, stats, bn_w = ctx.saved_tensors
        cfg = ctx.cfg
        gY = torch.zeros_like(Z) if gY is None else _f32c(gY)
        R, F = Z.shape
        dev = Z.device
        coef = torch.empty(3 * F + 2, device=dev)
        ws, wsb = workspace(2 * F, dev)
        gshift = None if gstats is None else _f32c(gstats[3 * F:])
        call("hgnn_bn_bwd_reduce", fptr(gY), fptr(Z), R, F, fptr(stats), fptr(bn_w),
             1 if ctx.training else 0, fptr(gshift), fptr(coef), ws, wsb, stream())
        gPre = torch.empty_like(Z)
        dbias = torch.empty(F, device=dev)
        call("hgnn_side_bwd_pre", fptr(gY), fptr(Z), R, F, fptr(coef), ctx.relu_from, fptr(gPre),
             fptr(dbias), ws, wsb, stream())
        gXs, gXc, dWa, dWb = _side_backward(cfg, gPre, Xs, Xc, Wa2, Wb2, ctx.needs_input_grad[0],
                                            ctx.needs_input_grad[1])
        Ha = Wa2.shape[0]
        return (gXs, gXc, dWa.view(ctx.wshapes[0]), dbias[:Ha], dWb.view(ctx.wshapes[1]), dbias[Ha:],
                coef[3 * F].reshape(bn_w.shape), coef[3 * F + 1].reshape(bn_w.shape),
                None, None, None, None)


def side_fits_fused(cfg, Xs, Xc, Fout):
    """Whether the fused side kernels (csrc/side.cu, csrc/engine.cu) can hold this side's resident weight
    block ``Cin x Fout`` in shared memory (``hgnn_lg_side_fits``: same budget arithmetic)."""
    from . import _lib
    return _lib.lib.hgnn_lg_side_fits(len(cfg.ops), Xs.shape[1], Xc.shape[1] if Xc is not None else 0, Fout) == 1


def side_update(Xs, Xc, Wa, ba, Wb, bb, bn_w, bn_b, cfg, relu_from, bn, training):
    """One layer side (see ``SideUpdate``).  Sides whose weight block exceeds the fused kernels' shared-memory
    budget (LGNN order 1 from h ~ 46: Cin x Fout = 10h x 2h floats) are composed instead from the
    stand-alone kernels - multi-operator gather (``Gmul``), a dense linear (there a real GEMM: cuBLAS),
    ReLU on the second branch, batch-norm on packed rows - with identical semantics
    (layers_mnb.py:200-212, :214-223)."""
    Ha, Hb = Wa.shape[0], Wb.shape[0]
    if side_fits_fused(cfg, Xs, Xc, Ha + Hb):
        return SideUpdate.apply(Xs, Xc, Wa, ba, Wb, bb, bn_w, bn_b, cfg, relu_from, bn, training)
    blocks = [Gmul.apply(Xs, cfg.ops, cfg.ops_T, cfg.R)]
    if Xc is not None and cfg.p is not None:
        blocks.append(Gmul.apply(Xc, [cfg.p.desc(False), cfg.p.desc(True)],
                                 [cfg.pt.desc(False), cfg.pt.desc(True)], cfg.R))
    x1 = torch.cat(blocks, 1) if len(blocks) > 1 else blocks[0]
    W = torch.cat([Wa.reshape(Ha, -1), Wb.reshape(Hb, -1)], 0)
    z = torch.nn.functional.linear(x1, W, torch.cat([ba, bb], 0))
    if relu_from < Ha + Hb:
        z = torch.cat([z[:, :relu_from], torch.relu(z[:, relu_from:])], 1)
    return BatchNormRows.apply(z, bn_w, bn_b, bn, training)


class Readout(torch.autograd.Function):
    """layer_last / layer_last_lg (layers_mnb.py:88-95, 379-388): fc over the gathered blocks, then
    the sum over ALL Nmax slots - padded slots contribute fc.bias each (SURVEY.md parity item 8)."""

    @staticmethod
    def forward(ctx, Xs, Xc, W, b, cfg, off, pad_count, bs):
        Xs = _f32c(Xs)
        Xc = _f32c(Xc) if Xc is not None else None
        W2 = _f32c(W).view(W.shape[0], -1)
        H = W2.shape[0]
        dev = Xs.device
        side, keep = _side_struct(cfg, Xs, Xc)
        Y1 = torch.empty(cfg.R, H, device=dev)
        call("hgnn_side_fwd", ctypes.byref(side), fptr(W2), fptr(b), H, None, None, 0, H, fptr(Y1),
             None, None, None, None, 0.0, None, None, 0, stream())
        out = torch.empty(bs, H, device=dev)
        call("hgnn_segment_sum", fptr(Y1), bs, H, iptr(off), fptr(pad_count), fptr(b), fptr(out), stream())
        ctx.save_for_backward(Xs, Xc, W2)
        ctx.cfg, ctx.off, ctx.bs, ctx.wshape = cfg, off, bs, W.shape
        ctx.n_slots = None
        ctx.pad_count = pad_count
        return out

    @staticmethod
    def backward(ctx, g):
        Xs, Xc, W2 = ctx.saved_tensors
        cfg = ctx.cfg
        g = _f32c(g)
        H = W2.shape[0]
        G = torch.empty(cfg.R, H, device=g.device)
        call("hgnn_segment_bcast", fptr(g), ctx.bs, H, iptr(ctx.off), fptr(G), stream())
        gXs, gXc, dW, _ = _side_backward(cfg, G, Xs, Xc, W2, None, ctx.needs_input_grad[0],
                                         ctx.needs_input_grad[1])
        # every one of the Nmax slots of graph b adds bias: d bias = sum_b Nmax * g[b]
        n_b = (ctx.off[1:] - ctx.off[:-1]).to(g.dtype) + ctx.pad_count
        dbias = (g * n_b.view(-1, 1)).sum(0)
        return gXs, gXc, dW.view(ctx.wshape), dbias, None, None, None, None


# --------------------------------------------------------------------------------------------
# CCN covariant contraction  (functions/contraction.py, functions/utils_ccn.py)
# --------------------------------------------------------------------------------------------


class Collapse6to3(torch.autograd.Function):
    """collapse6to3 on a general (C, n, n, n, n, n) tensor (contraction.py:106-121)."""

    @staticmethod
    def forward(ctx, F6):
        F6 = _f32c(F6)
        C, n = F6.shape[0], F6.shape[1]
        out = torch.empty(n, n, 18 * C, device=F6.device)
        call("hgnn_ccn2_collapse6to3", fptr(F6), C, n, fptr(out), stream())
        ctx.dims = (C, n)
        return out

    @staticmethod
    def backward(ctx, g):
        C, n = ctx.dims
        g = _f32c(g)
        gF = torch.empty((C,) + (n,) * 5, device=g.device)
        call("hgnn_ccn2_collapse6to3_bwd", fptr(g), C, n, fptr(gF), stream())
        return gF


class Ccn2Update(torch.autograd.Function):
    """One CCN-2D level for every vertex of the batch (utils_ccn.py:281-300), fused:
    promote -> 18 contractions -> Linear -> ReLU.  F tensors are packed (sum_v d_v^2, C)."""

    @staticmethod
    def forward(ctx, Fprev, W, b, st):
        Fprev, W, b = _f32c(Fprev), _f32c(W), _f32c(b)
        H, C = W.shape[0], Fprev.shape[1]
        Fnext = torch.empty(Fprev.shape[0], H, device=Fprev.device)
        call("hgnn_ccn2_update_fwd", st.V, st.nmax, iptr(st.nbr_ptr), iptr(st.nbr), st.f_off.data_ptr(),
             fptr(Fprev), C, fptr(W), fptr(b), H, fptr(Fnext), stream())
        ctx.save_for_backward(Fprev, W, Fnext)
        ctx.st = st
        return Fnext

    @staticmethod
    def backward(ctx, g):
        Fprev, W, Fnext = ctx.saved_tensors
        st = ctx.st
        g = _f32c(g)
        H, C = W.shape[0], Fprev.shape[1]
        gF = torch.empty_like(Fprev) if ctx.needs_input_grad[0] else None
        dW, db = torch.empty_like(W), torch.empty(H, device=W.device)
        ws, wsb = workspace(H * 18 * C + H, W.device)
        call("hgnn_ccn2_update_bwd", st.V, st.nmax, iptr(st.nbr_ptr), iptr(st.nbr), st.f_off.data_ptr(),
             fptr(Fprev), C, fptr(W), H, fptr(Fnext), fptr(g), fptr(gF), fptr(dW), fptr(db), ws, wsb, stream())
        return gF, dW, db, None


class Ccn1Update(torch.autograd.Function):
    """One CCN-1D level (utils_ccn.py:303-324).  F tensors are packed (sum_v d_v, C)."""

    @staticmethod
    def forward(ctx, Fprev, W, b, st):
        Fprev, W, b = _f32c(Fprev), _f32c(W), _f32c(b)
        H, C = W.shape[0], Fprev.shape[1]
        Fnext = torch.empty(Fprev.shape[0], H, device=Fprev.device)
        call("hgnn_ccn1_update_fwd", st.V, st.nmax, iptr(st.nbr_ptr), iptr(st.nbr), fptr(Fprev), C,
             fptr(W), fptr(b), H, fptr(Fnext), stream())
        ctx.save_for_backward(Fprev, W, Fnext)
        ctx.st = st
        return Fnext

    @staticmethod
    def backward(ctx, g):
        Fprev, W, Fnext = ctx.saved_tensors
        st = ctx.st
        g = _f32c(g)
        H, C = W.shape[0], Fprev.shape[1]
        gF = torch.empty_like(Fprev) if ctx.needs_input_grad[0] else None
        dW, db = torch.empty_like(W), torch.empty(H, device=W.device)
        ws, wsb = workspace(H * 2 * C + H, W.device)
        call("hgnn_ccn1_update_bwd", st.V, st.nmax, iptr(st.nbr_ptr), iptr(st.nbr), fptr(Fprev), C,
             fptr(W), H, fptr(Fnext), fptr(g), fptr(gF), fptr(dW), fptr(db), ws, wsb, stream())
        return gF, dW, db, None


class SegmentSum(torch.autograd.Function):
    """out[b] = sum of the packed rows of segment b (CCN readout, model_ccn.py:102, :61)."""

    @staticmethod
    def forward(ctx, Y, off, bs):
        Y = _f32c(Y)
        out = torch.empty(bs, Y.shape[1], device=Y.device)
        call("hgnn_segment_sum", fptr(Y), bs, Y.shape[1], iptr(off), None, None, fptr(out), stream())
        ctx.off, ctx.bs, ctx.R = off, bs, Y.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        G = torch.empty(ctx.R, g.shape[1], device=g.device)
        call("hgnn_segment_bcast", fptr(g), ctx.bs, g.shape[1], iptr(ctx.off), fptr(G), stream())
        return G, None, None
