"""CCN machinery - mirror of the reference's functions/utils_ccn.py (``CompnetUtils`` :28-324).

Same class and method names.  The reference keeps a Python list of per-vertex tensors and updates
them one vertex at a time with dense chi matmuls and the rank-6 contraction (:281-300); here the
per-vertex tensors are views into ONE packed buffer and a level update is a single kernel launch
over all vertices (of one graph, or of a whole batch of graphs via ``CcnStructure.from_graphs``).
chi is never materialised on the hot path (it is a partial permutation = an index map); the
``_get_chi`` / ``_promote`` helpers are kept for API parity and build small dense tensors on request.
"""
import numpy as np
import torch
import torch.nn.functional as Func

from .. import ops
from .._lib import require_cuda
from ..pack import SparseAdj
from .contraction import collapse6to3


class CcnStructure(object):
    """Receptive fields of a batch of graphs, resident on the GPU.

    ``nbr_ptr`` (V+1) / ``nbr``: sorted neighbour lists incl. self (utils_ccn.py:156-165), global
    vertex ids; ``f_off`` (V+1) int64: prefix sums of d_v^2 (2-D tiles); ``row_vertex2`` / ``row_vertex1``:
    owning vertex of every packed row (base features, :168-173 / :208-212); ``goff*``: per-graph
    segment offsets for the readout sums."""

    @classmethod
    def from_graphs(cls, adjs, device="cuda"):
        require_cuda()
        self = cls()
        nbrs, degs, vcount = [], [], []
        voff = 0
        for A in adjs:
            c, d, n = _graph_neighbours(A)
            nbrs.append(c + voff)
            degs.append(d)
            vcount.append(n)
            voff += n
        deg = np.concatenate(degs).astype(np.int64) if degs else np.zeros(0, np.int64)
        self.V = int(deg.shape[0])
        self.deg_host = deg
        self.nmax = int(deg.max()) if self.V else 1
        nbr_ptr = np.concatenate([[0], np.cumsum(deg)])
        f_off = np.concatenate([[0], np.cumsum(deg * deg)])
        self.nbr_ptr_host, self.f_off_host = nbr_ptr, f_off
        self.nbr_host = np.concatenate(nbrs) if nbrs else np.zeros(0, np.int64)
        dev = torch.device(device)
        self.device = dev
        gv = np.concatenate([[0], np.cumsum(vcount)]).astype(np.int64)
        self.n_graphs = len(adjs)
        self.vertex_off_host = gv
        # ONE pinned staging buffer and ONE host->device copy for the five int32 arrays and the two int64 arrays
        # (seven pageable copies cost ~0.15 ms of synchronous driver time per batch)
        vid = np.arange(self.V)
        i32 = [nbr_ptr, self.nbr_host, f_off[gv], nbr_ptr[gv]]
        i64 = [f_off, np.repeat(vid, deg * deg), np.repeat(vid, deg)]
        n32 = [int(x.shape[0]) for x in i32]
        n64 = [int(x.shape[0]) for x in i64]
        pad32 = (sum(n32) + 1) & ~1                      # keep the int64 part 8-byte aligned
        stage = torch.empty(pad32 * 4 + sum(n64) * 8, dtype=torch.uint8, pin_memory=True)
        h32 = stage[:pad32 * 4].view(torch.int32).numpy()
        h64 = stage[pad32 * 4:].view(torch.int64).numpy()
        pos = 0
        for x, n in zip(i32, n32):
            h32[pos:pos + n] = x
            pos += n
        pos = 0
        for x, n in zip(i64, n64):
            h64[pos:pos + n] = x
            pos += n
        buf = stage.to(dev, non_blocking=True)
        self._stage = stage                              # alive until the copy has run
        d32 = buf[:pad32 * 4].view(torch.int32).split_with_sizes(n32 + [pad32 - sum(n32)])
        d64 = buf[pad32 * 4:].view(torch.int64).split_with_sizes(n64)
        self.nbr_ptr, self.nbr, self.goff2, self.goff1 = d32[0], d32[1], d32[2], d32[3]
        self.f_off, self.row_vertex2, self.row_vertex1 = d64[0], d64[1], d64[2]
        return self


# per-graph neighbour lists (sorted, local vertex ids), degrees and vertex count, derived once per adjacency object - the
# same role as functions/batching._OPS_CACHE plays for the operator path (the reference recomputes them in every forward,
# utils_ccn.py:156-165)
_NBR_CACHE = {}


def _graph_neighbours(A):
    key = id(A)
    hit = _NBR_CACHE.get(key)
    if hit is not None and hit[0]() is A:
        return hit[1]
    if isinstance(A, SparseAdj):
        keep = A.vals > 0
        r, c, n = A.rows[keep], A.cols[keep], A.N
        order = np.lexsort((c, r))
        r, c = r[order], c[order]
    else:
        An = A.detach().cpu().numpy() if torch.is_tensor(A) else np.asarray(A)
        r, c = np.nonzero(An > 0)
        n = An.shape[0]
    out = (c.astype(np.int64), np.bincount(r, minlength=n), n)
    try:
        import weakref
        _NBR_CACHE[key] = (weakref.ref(A, lambda _r, k=key: _NBR_CACHE.pop(k, None)), out)
    except TypeError:
        pass
    return out


class PackedFeatures(list):
    """A list of per-vertex tensors (what the reference passes around) that are views into one
    packed buffer ``.packed`` of shape (sum_v d_v^2, C) (2-D) or (sum_v d_v, C) (1-D)."""

    def __init__(self, packed, st, order):
        self.packed, self.st, self.order = packed, st, order
        d = st.deg_host
        sizes = (d * d if order == 2 else d).tolist()
        C = packed.shape[1]
        views = torch.split(packed, sizes, 0)
        if order == 2:
            views = [v.view(int(k), int(k), C) for v, k in zip(views, d)]
        super(PackedFeatures, self).__init__(views)


class _ChiTable(object):
    """Lazy stand-in for the reference's ``self.chis`` list of lists (:125-145)."""

    def __init__(self, utils):
        self.u = utils

    def __getitem__(self, i):
        u = self.u

        class Row(object):
            def __getitem__(_s, j):
                n = u.st.V
                if j == n or j == -1:
                    return u._get_chi_root(i)
                return u._get_chi(i, j) if u._adjacent(i, j) else None
        return Row()


class CompnetUtils():
    def __init__(self, cudaflag=False):
        self.cudaflag = cudaflag
        self.outer_contract = self.python_contract
        self.st = None
        self._cache_key = None

    # ---- contraction of an explicit T (x) adj  (reference :37-45, :57-63) ----------------------
    def python_contract(self, T, adj):
        """T (n,n,n,C), adj (n,n) -> (n,n,18C): tensorprod then collapse6to3."""
        return collapse6to3(self.tensorprod(T.permute(3, 0, 1, 2), adj))

    def tensorprod(self, T, A):
        for i in range(A.dim()):
            T = torch.unsqueeze(T, T.dim())
        return T * A

    # ---- structure ------------------------------------------------------------------------------
    def _set_graph(self, A):
        require_cuda()
        key = (A.data_ptr(), A._version, tuple(A.shape)) if torch.is_tensor(A) else id(A)
        if key != self._cache_key:
            self.st = CcnStructure.from_graphs([A])
            self._cache_key, self._cache_A = key, A
        self.A = A
        st = self.st
        self.deg = torch.from_numpy(st.deg_host)
        self.neighbors = [torch.from_numpy(st.nbr_host[st.nbr_ptr_host[i]:st.nbr_ptr_host[i + 1]])
                          for i in range(st.V)]
        self.chis = _ChiTable(self)
        return st

    def _adjacent(self, i, j):
        return bool((self.neighbors[i] == j).any())

    def _get_chi(self, i, j):
        """chi (d_i, d_j): chi[k, l] = 1 iff neighbour k of i is neighbour l of j (:66-91)."""
        return (self.neighbors[i].view(-1, 1) == self.neighbors[j].view(1, -1)).float().to(self.st.device)

    def _get_chi_root(self, i):
        n = self.st.V
        chi = torch.zeros(n, int(self.deg[i]))
        chi[self.neighbors[i], torch.arange(int(self.deg[i]))] = 1
        return chi.to(self.st.device)

    def _register_chis(self, A):
        self._set_graph(A)
        return self.chis

    # ---- base features (reference :148-182, :185-222) -------------------------------------------
    def get_F0(self, X, A):
        st = self._set_graph(A)
        if not X.is_cuda:
            raise RuntimeError("hgnn_b200: CCN needs CUDA tensors (no CPU fallback)")
        return PackedFeatures(X.float().index_select(0, st.row_vertex2), st, 2)

    def get_F0_1D(self, X, A):
        st = self._set_graph(A)
        if not X.is_cuda:
            raise RuntimeError("hgnn_b200: CCN needs CUDA tensors (no CPU fallback)")
        return PackedFeatures(X.float().index_select(0, st.row_vertex1), st, 1)

    # ---- promotions, kept for API parity (reference :225-278); not used by update_F ----------
    def _promote(self, F_prev, i, j):
        chi = self._get_chi(i, j)
        return torch.matmul(chi, torch.matmul(F_prev[j].permute(2, 0, 1), chi.t())).permute(1, 2, 0)

    def _promote_1D(self, F_prev, i, j):
        return torch.matmul(self._get_chi(i, j), F_prev[j])

    def get_nbr_promotions(self, F_prev, i):
        return torch.stack([self._promote(F_prev, i, int(j)) for j in self.neighbors[i]], 0)

    def get_nbr_promotions_1D(self, F_prev, i):
        return torch.stack([self._promote_1D(F_prev, i, int(j)) for j in self.neighbors[i]], 0)

    # ---- level updates (reference :281-324): ONE kernel for all vertices -----------------------
    def _packed(self, F_prev, order):
        if isinstance(F_prev, PackedFeatures):
            return F_prev.packed, F_prev.st
        C = F_prev[0].shape[-1]
        return torch.cat([f.reshape(-1, C) for f in F_prev], 0), self.st

    def update_F(self, F_prev, W):
        packed, st = self._packed(F_prev, 2)
        assert len(F_prev) == st.V
        return PackedFeatures(ops.Ccn2Update.apply(packed, W.weight, W.bias, st), st, 2)

    def update_F_1D(self, F_prev, W):
        packed, st = self._packed(F_prev, 1)
        return PackedFeatures(ops.Ccn1Update.apply(packed, W.weight, W.bias, st), st, 1)
