"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A dense, CPU, fp32 restatement (numpy + torch-CPU) of the HGNN-2 hot path: operator construction,
batch padding, the multi-operator aggregation ("gmul"), the incidence products, the masked
batch-norm epilogue, the GNN / LGNN layer stacks and the CCN covariant contraction.  Every
function cites the reference file:line it follows (paths relative to the reference checkout).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module, and only as the checker or as the timed CPU baseline.  The
product package (``hgnn-2_b200/``) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference itself, executed in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (``tests/test_oracle_golden.py``),
and - where the reference checkout is present - against the imported reference directly
(``tests/test_oracle_vs_reference.py``).

Parameters are passed as plain dicts keyed by the reference's ``state_dict`` names
(``layer0.cv1.weight`` ...), so a reference model's weights drop straight in.
Gradients come from torch-CPU autograd over these dense formulas.
"""
import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# operator construction  (functions/operators.py:11-83)
# --------------------------------------------------------------------------------------------


def stored_edges(A):
    """The reference's ``edges`` table, functions/operators.py:46-66.

    The double loop walks i<j row-major and bumps ``e`` ONCE per undirected edge (:59), so with
    E undirected edges and M = nnz(A): row c<E holds forward edge c = (i_c, j_c, w_c), row E
    holds the reverse of the last edge, rows E+1..M-1 stay (0, 0, 0).
    """
    A = np.asarray(A, dtype=np.float32)
    M = int(np.count_nonzero(A))
    iu, ju = np.nonzero(np.triu(A != 0, 1))
    E = iu.shape[0]
    edges = np.zeros((M, 3), dtype=np.float32)
    if E:
        w = A[iu, ju]
        edges[:E, 0], edges[:E, 1], edges[:E, 2] = iu, ju, w
        edges[E] = (ju[-1], iu[-1], w[-1])
    return edges, iu, ju, E, M


def graph_operators(graph, J=1, dual=False):
    """functions/operators.py:11-83 (identical copy preprocessing/preprocessing.py:100-170).

    Vectorised and bug-compatible: weighted degree (:22-23), unclipped repeated squaring
    (:26-29, :78-81), the single ``e`` increment (:59) and ``AL[m1,m2] = w(m2)`` (:68-71).
    """
    V, A = graph
    A = np.asarray(A.detach().cpu() if torch.is_tensor(A) else A, dtype=np.float32)
    N = int(V.shape[0])
    W = np.zeros((N, N, J + 2), dtype=np.float32)
    W[:, :, 0] = np.eye(N, dtype=np.float32)
    W[:, :, 1] = np.diag(A.sum(axis=1, dtype=np.float32))
    W[:, :, 2] = A
    C = A.copy()
    for j in range(1, J):
        C = (torch.from_numpy(C) @ torch.from_numpy(C)).numpy()
        W[:, :, j + 2] = C
    if not dual:
        return torch.from_numpy(W)

    edges, iu, ju, E, M = stored_edges(A)
    Pm = np.zeros((N, M), dtype=np.float32)
    Pd = np.zeros((N, M), dtype=np.float32)
    if E:
        c = np.arange(E)
        # second half of iteration k writes column k+1 (:60-66) ...
        Pm[iu, c + 1] = 1
        Pm[ju, c + 1] = 1
        Pd[iu, c + 1] = -1
        Pd[ju, c + 1] = 1
        # ... and is then overwritten by the first half of iteration k+1 (:52-58)
        Pm[iu, c] = 1
        Pm[ju, c] = 1
        Pd[iu, c] = 1
        Pd[ju, c] = -1
    src, dst, w = edges[:, 0], edges[:, 1], edges[:, 2]
    link = (dst[:, None] == src[None, :]) & (src[:, None] != dst[None, :])      # :68-71
    AL = np.where(link, w[None, :], np.float32(0)).astype(np.float32)
    WL = np.zeros((M, M, J + 2), dtype=np.float32)
    WL[:, :, 0] = np.eye(M, dtype=np.float32)
    WL[:, :, 1] = np.diag(AL.sum(axis=1, dtype=np.float32))
    WL[:, :, 2] = AL
    CL = AL.copy()
    for j in range(1, J):
        CL = (torch.from_numpy(CL) @ torch.from_numpy(CL)).numpy()
        WL[:, :, j + 2] = CL
    return (torch.from_numpy(W), torch.from_numpy(WL), torch.from_numpy(Pm),
            torch.from_numpy(Pd))


# --------------------------------------------------------------------------------------------
# batching  (functions/batching.py:77-185)
# --------------------------------------------------------------------------------------------


def prepare_batch(batch, task, J=1):
    """functions/batching.py:77-185: zero-pad to (Nmax, Emax) and stack; XL = diag(WL[:,:,1])
    (:171); masks are 1 on the leading N_i x N_i / E_i x E_i blocks (:182-183); E_i = nnz(A_i)
    (:103) so phantom line-graph columns count as real."""
    bs = len(batch)
    F = batch[0][0].shape[1]
    N_batch = torch.tensor([b[0].shape[0] for b in batch], dtype=torch.int64)
    E_batch = torch.tensor([int(torch.count_nonzero(b[1])) for b in batch], dtype=torch.int64)
    Nmax, Emax = int(N_batch.max()), int(E_batch.max())
    K = J + 2
    X = torch.zeros(bs, F, Nmax)
    W = torch.zeros(bs, Nmax, Nmax, K)
    T = torch.zeros(bs, 1)
    XL = torch.zeros(bs, 1, Emax)
    WL = torch.zeros(bs, Emax, Emax, K)
    Pm = torch.zeros(bs, Nmax, Emax)
    Pd = torch.zeros(bs, Nmax, Emax)
    mask = torch.zeros(bs, Nmax, Nmax)
    mask_lg = torch.zeros(bs, Emax, Emax)
    for i, (x, A, t, w, wl, pm, pd) in enumerate(batch):
        n, e = int(N_batch[i]), int(E_batch[i])
        X[i, :, :n] = x.t()
        W[i, :n, :n] = w
        WL[i, :e, :e] = wl
        Pm[i, :n, :e] = pm
        Pd[i, :n, :e] = pd
        XL[i, 0, :e] = torch.diagonal(wl[:, :, 1])
        T[i, 0] = t[task]
        mask[i, :n, :n] = 1
        mask_lg[i, :e, :e] = 1
    return X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch


# --------------------------------------------------------------------------------------------
# gmul / pmul  (models/layers/layers_mnb.py:391-434 == functions/utils.py:24-81)
# --------------------------------------------------------------------------------------------


def graph_op(W, X):
    """layers_mnb.py:395-411: out[b, j*F+f, v] = sum_u W[b,v,u,j] X[b,f,u], one mm per (b, j)
    exactly like the reference's double loop (this is also what the CPU baseline times)."""
    bs, N, _, K = W.shape
    F = X.shape[1]
    blocks = []
    for b in range(bs):
        xb = X[b].t()
        blocks.append(torch.cat([torch.mm(W[b, :, :, j], xb).t() for j in range(K)], 0))
    return torch.stack(blocks, 0)


def pmul(P, X):
    """layers_mnb.py:418-434: out[b,f,v] = sum_e P[b,v,e] X[b,f,e], one mm per graph."""
    return torch.stack([torch.mm(P[b], X[b].t()).t() for b in range(P.shape[0])], 0)


# --------------------------------------------------------------------------------------------
# masked batch norm  (models/layers/batch_normalization.py:23-108)
# --------------------------------------------------------------------------------------------


def bn_stats(H, N_batch, mask):
    """batch_normalization.py:65-93: zero the padded slots, per-feature mean over the sum(N_b)
    real slots, biased variance with 1e-5 added inside the square root."""
    keep = mask[:, :, 0].unsqueeze(1)
    Hm = H * keep
    n = float(N_batch.sum().item())
    mean = Hm.sum(dim=(0, 2)) / n
    dev = (Hm - mean.view(1, -1, 1)) ** 2 * keep
    var = 1e-5 + dev.sum(dim=(0, 2)) / n
    return Hm, mean, var ** 0.5


def bn_forward(H, N_batch, mask, weight, bias, running=None):
    """batch_normalization.py:34-43: train mode (running=None) normalises EVERY slot, padded ones
    included, with batch stats; eval mode uses (running_mean, running_std).  Scalar affine.
    Returns (out, mean, std)."""
    if running is None:
        Hm, mean, std = bn_stats(H, N_batch, mask)
    else:
        Hm = H * mask[:, :, 0].unsqueeze(1)
        mean, std = running
    out = (Hm - mean.view(1, -1, 1)) / std.view(1, -1, 1)
    return weight * out + bias, mean, std


# --------------------------------------------------------------------------------------------
# layers  (models/layers/layers_mnb.py)
# --------------------------------------------------------------------------------------------


def _conv(p, name, x):
    # Conv1d(kernel=1): weight (Fout, Cin, 1), bias (Fout,)
    return torch.einsum("oc,bcn->bon", p[name + ".weight"][:, :, 0], x) + p[name + ".bias"].view(1, -1, 1)


def _bn(p, name, H, n_batch, mask, stats):
    out, mean, std = bn_forward(H, n_batch, mask, p[name + ".weight"], p[name + ".bias"],
                                None if stats is None or stats.get("eval") is None
                                else stats["eval"][name])
    if stats is not None:
        stats.setdefault("batch", {})[name] = (mean.detach(), std.detach())
    return out


def layer_simple(p, pre, state, N_batch, mask, stats=None):
    """layers_mnb.py:52-69: both conv branches are ReLU'd (:61,:64); cat(cv2-branch, cv1-branch)."""
    X, W = state
    x1 = graph_op(W, X)
    z1 = torch.relu(_conv(p, pre + "cv1", x1))
    yl1 = torch.relu(_conv(p, pre + "cv2", x1))
    return _bn(p, pre + "bn1", torch.cat((yl1, z1), 1), N_batch, mask, stats), W


def layer_last(p, pre, state):
    """layers_mnb.py:88-95: fc then sum over ALL Nmax slots (padded slots contribute fc.bias)."""
    X, W = state
    return _conv(p, pre + "fc", graph_op(W, X)).sum(dim=2)


def layer_with_lg(order, p, pre, state, N_batch, mask, E_batch, mask_lg, stats=None):
    """layers_mnb.py:189-225 (order 1), :256-290 (order 2), :322-358 (order 3)."""
    X, XL, W, WL, Pm, Pd = state
    xa1 = graph_op(W, X)
    xda1 = graph_op(WL, XL)
    PmT, PdT = Pm.transpose(2, 1), Pd.transpose(2, 1)

    def node_update(edge_state):
        x1 = torch.cat((xa1, pmul(Pm, edge_state), pmul(Pd, edge_state)), 1)
        zb1 = torch.cat((_conv(p, pre + "cv2", x1), torch.relu(_conv(p, pre + "cv1", x1))), 1)
        return _bn(p, pre + "bn1", zb1, N_batch, mask, stats)

    def edge_update(node_state):
        xd1 = torch.cat((xda1, pmul(PmT, node_state), pmul(PdT, node_state)), 1)
        zdb1 = torch.cat((_conv(p, pre + "cv4", xd1), torch.relu(_conv(p, pre + "cv3", xd1))), 1)
        return _bn(p, pre + "bn2", zdb1, E_batch, mask_lg, stats)

    if order == 1:      # edges see the NEW node state
        zbn1 = node_update(XL)
        zdbn1 = edge_update(zbn1)
    elif order == 2:    # nodes see the NEW edge state
        zdbn1 = edge_update(X)
        zbn1 = node_update(zdbn1)
    else:               # both from the old states
        zbn1 = node_update(XL)
        zdbn1 = edge_update(X)
    return zbn1, zdbn1, W, WL, Pm, Pd


def layer_last_lg(p, pre, state):
    """layers_mnb.py:379-388."""
    X, XL, W, WL, Pm, Pd = state
    x1 = torch.cat((graph_op(W, X), pmul(Pm, XL), pmul(Pd, XL)), 1)
    return _conv(p, pre + "fc", x1).sum(dim=2)


def gnn_simple_forward(p, n_layers, state, N_batch, mask, stats=None):
    """models/gnns/model_mnb.py:58-66."""
    cur = layer_simple(p, "layer0.", state, N_batch, mask, stats)
    for i in range(n_layers - 2):
        cur = layer_simple(p, "layer%d." % (i + 1), cur, N_batch, mask, stats)
    return layer_last(p, "layerlast.", cur)


def gnn_lg_forward(p, n_layers, order, state, N_batch, mask, E_batch, mask_lg, stats=None):
    """models/gnns/model_mnb.py:124-129."""
    cur = layer_with_lg(order, p, "layer0.", state, N_batch, mask, E_batch, mask_lg, stats)
    for i in range(n_layers - 2):
        cur = layer_with_lg(order, p, "layer%d." % (i + 1), cur, N_batch, mask, E_batch, mask_lg,
                            stats)
    return layer_last_lg(p, "layerlast.", cur)


def init_gnn_params(kind, n_features, n_layers, dim_input, dim_output=1, J=1, order=1, seed=0):
    """Parameter dict with the reference's names/shapes/init scale (layers_mnb.py:36-50,172-187,
    239-254,305-320,371-377; batch_normalization.py:26-29).  RNG order is NOT the reference's;
    parity tests copy weights instead of replaying RNG (SURVEY.md parity item 14)."""
    g = torch.Generator().manual_seed(seed)
    K, h = J + 2, n_features

    def nrm(*shape):
        return torch.randn(*shape, generator=g) * 0.1

    p = {}
    for li in range(n_layers - 1):
        pre = "layer%d." % li
        if kind == "simple":
            fn = dim_input if li == 0 else 2 * h
            convs = {"cv1": K * fn, "cv2": K * fn}
            bns = ["bn1"]
        else:
            fn, fe = (dim_input, 1) if li == 0 else (2 * h, 2 * h)
            if order == 1:
                convs = {"cv1": K * fn + 2 * fe, "cv2": K * fn + 2 * fe,
                         "cv3": K * fe + 4 * h, "cv4": K * fe + 4 * h}
            elif order == 2:
                convs = {"cv1": K * fn + 4 * h, "cv2": K * fn + 4 * h,
                         "cv3": K * fe + 2 * fn, "cv4": K * fe + 2 * fn}
            else:
                convs = {"cv1": K * fn + 2 * fe, "cv2": K * fn + 2 * fe,
                         "cv3": K * fe + 2 * fn, "cv4": K * fe + 2 * fn}
            bns = ["bn1", "bn2"]
        for name, cin in convs.items():
            p[pre + name + ".weight"] = nrm(h, cin, 1)
            p[pre + name + ".bias"] = nrm(h)
        for name in bns:
            p[pre + name + ".weight"] = nrm(1).reshape(())
            p[pre + name + ".bias"] = nrm(1).reshape(())
    cin = K * 2 * h if kind == "simple" else (K + 2) * 2 * h
    p["layerlast.fc.weight"] = nrm(dim_output, cin, 1)
    p["layerlast.fc.bias"] = nrm(dim_output)
    return p


# --------------------------------------------------------------------------------------------
# CCN second-order contraction  (functions/contraction.py, functions/utils_ccn.py)
# --------------------------------------------------------------------------------------------

_P111 = [(0, 1, 2, 3, 4), (0, 3, 1, 2, 4), (1, 2, 0, 3, 4), (1, 3, 0, 2, 4), (3, 4, 0, 1, 2)]
_P12 = [(0, 1, 4, 2, 3)] + [(0, 1, 2, 3, 4)] * 9      # contraction.py:69-80 (identity repeated 9x)
_P3 = [(0, 3, 1, 2, 4), (1, 3, 0, 2, 4), (3, 4, 0, 1, 2)]


def collapse6to3(F):
    """functions/contraction.py:106-121.  F is (C, n, n, n, n, n); returns (n, n, 18*C) with
    contraction k, channel c at column k*C+c (:118).  Each case permutes the five spatial axes,
    optionally masks with a planar (:38-39) or cubic (:35-37) diagonal aligned to permuted axes
    (3,4) / (2,3,4), and sums permuted axes 2,3,4 (:21-26)."""
    C, n = F.shape[0], F.shape[1]
    G = F.permute(1, 2, 3, 4, 5, 0)
    eye = torch.eye(n, dtype=F.dtype)
    planar = eye.view(1, 1, 1, n, n, 1)
    cubic = (eye.unsqueeze(2) * eye).view(1, 1, n, n, n, 1)
    out = []
    for perm in _P111:
        out.append(G.permute(*perm, 5).sum(dim=(2, 3, 4)))
    for perm in _P12:
        out.append((G.permute(*perm, 5) * planar).sum(dim=(2, 3, 4)))
    for perm in _P3:
        out.append((G.permute(*perm, 5) * cubic).sum(dim=(2, 3, 4)))
    return torch.cat(out, 2)


def outer_contract(T, adj):
    """functions/utils_ccn.py:37-45 + :57-63: H = T (x) adj then collapse6to3.  T is (n,n,n,C)."""
    Tc = T.permute(3, 0, 1, 2)
    H = Tc[:, :, :, :, None, None] * adj
    return collapse6to3(H)


def outer_contract_closed_form(T, adj):
    """The 18 blocks of ``outer_contract`` in closed form (verified
    there against the reference in fp64).  Used to pin the formulas the CUDA kernel implements."""
    s = adj.sum()
    r = adj.sum(dim=1)
    tr = torch.diagonal(adj).sum()
    Sc = T.sum(dim=2)            # [a,b,:]
    Sa = T.sum(dim=0)            # [b,c,:]
    Sbc = T.sum(dim=(1, 2))      # [a,:]
    Sac = T.sum(dim=(0, 2))      # [b,:]
    Sall = T.sum(dim=(0, 1, 2))
    n = T.shape[0]
    idx = torch.arange(n)
    blocks = [
        Sc * s,
        Sbc[:, None, :] * r[None, :, None],
        Sa * s,
        Sac[:, None, :] * r[None, :, None],
        adj[:, :, None] * Sall,
        torch.einsum("abcf,c->abf", T, r),
    ] + [Sc * tr] * 9 + [
        torch.einsum("abf,db->adf", T[:, idx, idx, :], adj),
        torch.einsum("abf,da->bdf", T[idx, :, idx, :], adj),   # T[a,b,a] (advanced dims lead)
        adj[:, :, None] * T[idx, idx, idx, :].sum(dim=0),
    ]
    return torch.cat(blocks, 2)


def receptive_fields(A):
    """functions/utils_ccn.py:156-165: neighbours of v in A (ascending; includes v iff A[v,v]>0...
    the reference takes the first deg(v) entries of nonzero(A[v]) with deg = #(A>0))."""
    A = A.detach()
    return [torch.nonzero(A[i] > 0).flatten() for i in range(A.shape[0])]


def promote(F_j, nbr_i, nbr_j):
    """functions/utils_ccn.py:225-239 with chi from :66-91: chi[k, l] = 1 iff nbr_i[k] == nbr_j[l];
    returns chi F_j chi^T, shape (d_i, d_i, C)."""
    chi = (nbr_i.view(-1, 1) == nbr_j.view(1, -1)).to(F_j.dtype)
    return torch.einsum("kp,pqc,lq->klc", chi, F_j, chi)


def ccn2_forward(p, n_layers, X, A):
    """models/compnets/model_ccn.py:93-105 + functions/utils_ccn.py:148-182,255-300.
    ``A`` must already contain self-loops (scripts/train_ccn.py:36)."""
    nbrs = receptive_fields(A)
    n = A.shape[0]
    cur = [X[i].view(1, 1, -1).expand(len(nbrs[i]), len(nbrs[i]), -1) for i in range(n)]
    levels = [cur]
    for lvl in range(n_layers):
        w, b = p["w%d.weight" % (lvl + 1)], p["w%d.bias" % (lvl + 1)]
        new = []
        for i in range(n):
            T = torch.stack([promote(cur[int(j)], nbrs[i], nbrs[int(j)]) for j in nbrs[i]], 0)
            adj_i = (nbrs[i].view(-1, 1) == nbrs[i].view(1, -1)).to(X.dtype)   # chis[i][i], :293
            new.append(torch.relu(outer_contract(T, adj_i) @ w.t() + b))
        cur = new
        levels.append(cur)
    feat = torch.cat([sum(v.sum(dim=(0, 1)) for v in lvl) for lvl in levels], 0)
    return p["fc.weight"] @ feat + p["fc.bias"]


def ccn1_forward(p, n_layers, X, A):
    """models/compnets/model_ccn.py:41-64 + functions/utils_ccn.py:185-222,242-252,269-278,303-324."""
    nbrs = receptive_fields(A)
    n = A.shape[0]
    cur = [X[i].view(1, -1).expand(len(nbrs[i]), -1) for i in range(n)]
    levels = [cur]
    for lvl in range(n_layers):
        w, b = p["w%d.weight" % (lvl + 1)], p["w%d.bias" % (lvl + 1)]
        new = []
        for i in range(n):
            T = torch.stack([(nbrs[i].view(-1, 1) == nbrs[int(j)].view(1, -1)).to(X.dtype)
                             @ cur[int(j)] for j in nbrs[i]], 0)
            new.append(torch.relu(torch.cat([T.sum(0), T.sum(1)], 1) @ w.t() + b))
        cur = new
        levels.append(cur)
    feat = torch.cat([sum(v.sum(0) for v in lvl) for lvl in levels], 0)
    return p["fc.weight"] @ feat + p["fc.bias"]


def init_ccn_params(order, input_feats, n_outputs, hidden, n_layers, seed=0):
    """model_ccn.py:27-39 (1-D) / :79-91 (2-D) shapes."""
    g = torch.Generator().manual_seed(seed)
    nc = 2 if order == 1 else 18
    p = {}
    for i in range(n_layers):
        cin = (input_feats if i == 0 else hidden) * nc
        p["w%d.weight" % (i + 1)] = torch.randn(hidden, cin, generator=g) * 0.1
        p["w%d.bias" % (i + 1)] = torch.randn(hidden, generator=g) * 0.1
    p["fc.weight"] = torch.randn(n_outputs, n_layers * hidden + input_feats, generator=g) * 0.5
    p["fc.bias"] = torch.randn(n_outputs, generator=g) * 0.1
    return p
