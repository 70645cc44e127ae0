"""Host side of the operator pack: per-graph sparse operators and their block-diagonal batch.

The reference materialises dense per-graph tensors ``W (N,N,J+2)``, ``WL (M,M,J+2)``,
``Pm, Pd (N,M)`` (functions/operators.py:11-83) and zero-pads them per batch
(functions/batching.py:77-185).  Here the same operators are held as CSR index/value arrays -
exactly the non-zeros the reference would produce, quirks included:

* ``W[:,:,1]`` is the weighted degree, A keeps its weights (operators.py:22-25);
* the line-graph enumeration bumps ``e`` once per undirected edge (operators.py:59), so of the
  M = nnz(A) line-graph nodes only columns 0..E hold an edge (column E = reverse of the last
  edge) and columns E+1..M-1 are phantom ``(0,0,0)`` rows that still link to every stored edge
  leaving node 0 (operators.py:68-71);
* ``AL[m1,m2] = w(m2)``; ``Pm``/``Pd`` columns are the union of edges c and c-1 with the
  later write winning (operators.py:52-66).

Powers ``A^(2^j)`` / ``AL^(2^j)`` are NOT built here: they are squared on the GPU by the SpGEMM
kernel at batch-pack time (``pack.py``).  This module is pure index bookkeeping (numpy, O(nnz));
all arithmetic on features happens in the CUDA kernels.
"""
import numpy as np

I32 = np.int32
F32 = np.float32


def _csr_from_sorted_coo(n_rows, rows, cols, *vals):
    """COO sorted by (row, col) -> (rowptr, col, *vals)."""
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return (np.cumsum(rowptr).astype(I32), cols.astype(I32)) + tuple(v.astype(F32) for v in vals)


def _sort_coo(rows, cols, *vals):
    order = np.lexsort((cols, rows))
    return (rows[order], cols[order]) + tuple(v[order] for v in vals)


def split_runs(n_rows, rowptr, col, val, min_run=64):
    """Run-length split of a CSR matrix: every maximal run of >= ``min_run`` consecutive columns
    with one value inside a row is removed from the CSR and recorded as a *range entry*
    ``(row, [lo, hi), val)`` meaning ``val * sum_{c in [lo,hi)} x[c]``.

    The reference's phantom line-graph nodes (operators.py:59,68-71) make the transposed operator
    ``AL^T`` exactly this shape: the ~deg(0) rows of the edges leaving node 0 each hold one entry per
    phantom node, i.e. a run over the contiguous phantom range - 85 % of nnz(AL) at N=1000.  The
    kernels evaluate a range sum once per CTA instead of once per entry.

    Returns (rowptr2, col2, val2, rng_rowptr, rng_id, rng_val, rng_lo, rng_hi); ``rng_id`` indexes
    the table of distinct ranges (rng_lo, rng_hi)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    nnz = col.shape[0]
    empty_i, empty_f = np.zeros(0, I32), np.zeros(0, F32)
    if nnz == 0:
        z = np.zeros(n_rows + 1, I32)
        return z, empty_i, empty_f, z.copy(), empty_i, empty_f, empty_i, empty_i
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(rowptr))
    c = col.astype(np.int64)
    brk = np.ones(nnz, dtype=bool)
    brk[1:] = (rows[1:] != rows[:-1]) | (c[1:] != c[:-1] + 1) | (val[1:] != val[:-1])
    run_id = np.cumsum(brk) - 1
    run_start = np.nonzero(brk)[0]
    run_len = np.diff(np.append(run_start, nnz))
    long_run = run_len >= min_run
    keep = ~long_run[run_id]
    rp2 = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rp2, rows[keep] + 1, 1)
    ls = run_start[long_run]
    r_rows, r_lo, r_hi, r_val = rows[ls], c[ls], c[ls] + run_len[long_run], val[ls]
    uniq, inv = (np.unique(np.stack([r_lo, r_hi], 1), axis=0, return_inverse=True)
                 if ls.size else (np.zeros((0, 2), np.int64), np.zeros(0, np.int64)))
    rrp = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rrp, r_rows + 1, 1)
    return (np.cumsum(rp2).astype(I32), col[keep].astype(I32), val[keep].astype(F32),
            np.cumsum(rrp).astype(I32), inv.reshape(-1).astype(I32), r_val.astype(F32),
            uniq[:, 0].astype(I32), uniq[:, 1].astype(I32))


def _numpy_blob_alloc(nbytes):
    return np.empty(nbytes, dtype=np.uint8)


# where graph blobs live: plain numpy memory by default; ``pack.py`` installs a pinned-slab allocator
# when a CUDA device is present, so that a batch is copied host->device straight out of the dataset
blob_alloc = _numpy_blob_alloc

BLOB_MAGIC = 0x48474e4e424c4f42
BLOB_FIELDS = ("deg", "a_rowptr", "a_col", "a_val", "at_rowptr", "at_col", "at_val",
               "dl", "b_rowptr", "b_col", "b_val",
               "p_rowptr", "p_col", "p_pm", "p_pd", "pt_rowptr", "pt_col", "pt_pm", "pt_pd",
               "bts_rowptr", "bts_col", "bts_val", "bts_rng_rowptr", "bts_rng_id", "bts_rng_val",
               "bts_rng_lo", "bts_rng_hi",
               "btc_rowptr", "btc_col", "btc_val", "ew", "erow",
               "bt_rowptr", "bt_col", "bt_val")     # last: batches that skip the full bt copy a prefix
N_PRIMAL_FIELDS = 7


class GraphOps(object):
    """Sparse twin of ``graph_operators([V, A], J, dual=True)`` for ONE graph (local indices).

    Attributes (numpy): ``N, M, E``; A as CSR ``a_*`` and its transpose ``at_*``; weighted degree
    ``deg``; the line-graph operator ``b_*`` / ``bt_*`` with ``dl`` = its row sums (this is also
    the initial edge feature XL, batching.py:171); incidence ``p_*`` (rows = nodes) and ``pt_*``
    (rows = line-graph nodes), each with the two value arrays ``pm`` and ``pd`` on one pattern.
    """

    __slots__ = ("N", "M", "E", "a_rowptr", "a_col", "a_val", "at_rowptr", "at_col", "at_val",
                 "deg", "b_rowptr", "b_col", "b_val", "bt_rowptr", "bt_col", "bt_val", "dl",
                 "p_rowptr", "p_col", "p_pm", "p_pd", "pt_rowptr", "pt_col", "pt_pm", "pt_pd",
                 "bts_rowptr", "bts_col", "bts_val", "bts_rng_rowptr", "bts_rng_id", "bts_rng_val",
                 "bts_rng_lo", "bts_rng_hi", "btc_rowptr", "btc_col", "btc_val", "ew", "erow",
                 "dual", "_blob", "_blob_ptr")

    def blob_ptr(self):
        """Address of this graph's contiguous host blob (built on first use): int64 header
        ``[magic, N, M, E, n_fields, (byte offset, length) x n_fields]`` followed by the arrays of
        ``BLOB_FIELDS`` (16-byte aligned).  The attribute arrays are re-pointed at the blob, so it is
        the single copy of the graph.  ``hgnn_host_pack_fill`` (csrc/hostpack.cu) concatenates a
        batch straight from these blobs."""
        ptr = getattr(self, "_blob_ptr", None)
        if ptr is not None:
            return ptr
        fields = BLOB_FIELDS if self.dual else BLOB_FIELDS[:N_PRIMAL_FIELDS]
        arrs = [np.ascontiguousarray(getattr(self, f)) for f in fields]
        for f, a in zip(fields, arrs):
            if a.dtype.itemsize != 4:
                raise TypeError("GraphOps.%s must be a 4-byte array, got %s" % (f, a.dtype))
        head = 8 * (5 + 2 * len(fields))
        pos = (head + 15) & ~15
        table = []
        for a in arrs:
            table += [pos, a.shape[0]]
            pos += (a.nbytes + 15) & ~15
        blob = blob_alloc(pos)
        blob[:head].view(np.int64)[:] = [BLOB_MAGIC, self.N, self.M, self.E, len(fields)] + table
        for f, a, o in zip(fields, arrs, table[0::2]):
            view = blob[o:o + a.nbytes].view(a.dtype)
            view[:] = a
            setattr(self, f, view)
        self._blob = blob
        self._blob_ptr = blob.ctypes.data
        return self._blob_ptr

    def __getstate__(self):       # pickled datasets (functions/data_generator.py) carry plain arrays
        return {s: getattr(self, s) for s in self.__slots__ if not s.startswith("_blob") and hasattr(self, s)}

    def __setstate__(self, state):
        for k, v in state.items():
            setattr(self, k, v)
        if getattr(self, "dual", False) and "ew" not in state:     # pickled before the collapsed fields existed
            self._build_collapsed()

    def _build_collapsed(self):
        """Duplicate-row compression of the line graph (used by the persistent engine kernels).

        The reference's enumeration (operators.py:59) leaves line-graph rows E+1..M-1 as identical
        phantom ``(0,0,0)`` entries: same row of ``AL`` (the edges leaving node 0, :68-71), empty
        ``Pm``/``Pd`` columns, same degree - and no operator has an entry in their COLUMNS.  With the
        degree as the edge feature (batching.py:171) they therefore carry identical values in every
        layer, forward and backward, and influence the result only through sums over rows (batch-norm
        statistics, weight gradients).  So one representative (row E+1) is computed with multiplicity
        ``mu = M-E-1`` and the other phantom rows are skipped:

        * ``erow``  - the active rows 0..E+1 (all rows when there is no phantom block);
        * ``ew``    - per-row weight in every sum over rows: 1, ``mu`` for the representative; the skipped
                      rows hold ``-(distance to their representative)`` (weight 0; the kernels use it to
                      copy the representative's value where a full tensor is needed);
        * ``btc_*`` - the transposed operator restricted to active source rows, the representative's
                      entries scaled by ``mu``:  ``btc[r, k] = ew[k] * AL[k, r]``."""
        M, E = self.M, self.E
        n_ph = M - E - 1
        ew = np.ones(M, dtype=F32)
        if n_ph >= 2:
            ew[E + 1] = n_ph
            ew[E + 2:] = -np.arange(1, n_ph, dtype=F32)     # skipped copies: -(distance to the representative)
            n_act = E + 2
        else:
            n_act = M
        self.ew = ew
        self.erow = np.arange(n_act, dtype=I32)
        rp = np.asarray(self.b_rowptr, dtype=np.int64)
        m1 = np.repeat(np.arange(M, dtype=np.int64), np.diff(rp))
        keep = m1 < n_act
        m1, m2 = m1[keep], np.asarray(self.b_col, dtype=np.int64)[keep]
        v = (np.asarray(self.b_val, dtype=F32)[keep] * np.maximum(ew[m1], 0)).astype(F32)
        t1, t2, tv = _sort_coo(m2, m1, v)
        self.btc_rowptr, self.btc_col, self.btc_val = _csr_from_sorted_coo(M, t1, t2, tv)

    @classmethod
    def from_dense(cls, A, dual=True):
        A = np.asarray(A, dtype=F32)
        r, c = np.nonzero(A)                       # row-major order
        return cls.from_coo(A.shape[0], r, c, A[r, c], dual=dual, presorted=True)

    @classmethod
    def from_coo(cls, N, rows, cols, vals, dual=True, presorted=False):
        """(rows, cols, vals): every stored entry of A, both directions of each edge."""
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        vals = np.asarray(vals, dtype=F32)
        keep = vals != 0
        if not keep.all():
            rows, cols, vals = rows[keep], cols[keep], vals[keep]
        if not presorted:
            rows, cols, vals = _sort_coo(rows, cols, vals)
        self = cls()
        self.N, self.dual = int(N), bool(dual)
        self.a_rowptr, self.a_col, self.a_val = _csr_from_sorted_coo(N, rows, cols, vals)
        tr, tc, tv = _sort_coo(cols, rows, vals)
        self.at_rowptr, self.at_col, self.at_val = _csr_from_sorted_coo(N, tr, tc, tv)
        deg = np.zeros(N, dtype=F32)
        np.add.at(deg, rows, vals)       # small exact values: order-independent
        self.deg = deg
        self.M = int(vals.shape[0])
        if not dual:
            self.E = int(np.count_nonzero(cols > rows))
            return self
        self._build_line_graph(rows, cols, vals)
        return self

    # ------------------------------------------------------------------------------------
    def _build_line_graph(self, rows, cols, vals):
        N, M = self.N, self.M
        fwd = cols > rows                                   # operators.py:49-51, i<j row-major
        iu, ju, wu = rows[fwd], cols[fwd], vals[fwd]
        E = self.E = int(iu.shape[0])
        if E and M <= E:
            raise ValueError("adjacency must be symmetric: the reference writes line-graph "
                             "column E, which needs M=nnz(A) > E (operators.py:60-66)")
        src = np.zeros(M, dtype=np.int64)
        dst = np.zeros(M, dtype=np.int64)
        w = np.zeros(M, dtype=F32)
        if E:
            src[:E], dst[:E], w[:E] = iu, ju, wu
            src[E], dst[E], w[E] = ju[-1], iu[-1], wu[-1]
        # ---- AL[m1, m2] = w(m2) iff dst(m1) == src(m2) and src(m1) != dst(m2)  (:68-71)
        fstart = np.zeros(N + 1, dtype=np.int64)            # forward edges are grouped by source
        np.add.at(fstart, iu + 1, 1)
        fstart = np.cumsum(fstart)
        cnt = fstart[dst + 1] - fstart[dst]                 # candidates per line-graph node m1
        m1 = np.repeat(np.arange(M, dtype=np.int64), cnt)
        within = np.arange(m1.shape[0], dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        m2 = fstart[dst[m1]] + within
        if E:                                               # the one stored reverse edge
            extra = np.nonzero(dst == src[E])[0]
            m1 = np.concatenate([m1, extra])
            m2 = np.concatenate([m2, np.full(extra.shape[0], E, dtype=np.int64)])
        ok = src[m1] != dst[m2]
        m1, m2 = m1[ok], m2[ok]
        bv = w[m2]
        m1, m2, bv = _sort_coo(m1, m2, bv)
        self.b_rowptr, self.b_col, self.b_val = _csr_from_sorted_coo(M, m1, m2, bv)
        t1, t2, tv = _sort_coo(m2, m1, bv)
        self.bt_rowptr, self.bt_col, self.bt_val = _csr_from_sorted_coo(M, t1, t2, tv)
        # run-length split of the transposed operator for the engine kernels (phantom ranges)
        (self.bts_rowptr, self.bts_col, self.bts_val, self.bts_rng_rowptr, self.bts_rng_id,
         self.bts_rng_val, self.bts_rng_lo, self.bts_rng_hi) = split_runs(M, self.bt_rowptr, self.bt_col, self.bt_val)
        dl = np.zeros(M, dtype=F32)
        np.add.at(dl, m1, bv)
        self.dl = dl
        self._build_collapsed()
        # ---- Pm / Pd  (:52-66): column c gets edge c (+1 at i, -1 at j) written AFTER edge c-1
        #      (-1 at i, +1 at j); Pm is 1 on the union.
        c = np.arange(E, dtype=np.int64)
        node = np.concatenate([iu, ju, iu, ju])
        colm = np.concatenate([c + 1, c + 1, c, c])
        pdv = np.concatenate([-np.ones(E), np.ones(E), np.ones(E), -np.ones(E)]).astype(F32)
        prio = np.concatenate([np.zeros(2 * E), np.ones(2 * E)])      # later write wins
        order = np.lexsort((prio, colm, node))
        node, colm, pdv = node[order], colm[order], pdv[order]
        last = np.ones(node.shape[0], dtype=bool)
        if node.shape[0] > 1:
            last[:-1] = (node[1:] != node[:-1]) | (colm[1:] != colm[:-1])
        node, colm, pdv = node[last], colm[last], pdv[last]
        pmv = np.ones(node.shape[0], dtype=F32)
        self.p_rowptr, self.p_col, self.p_pm, self.p_pd = _csr_from_sorted_coo(N, node, colm, pmv, pdv)
        tc, tn, tpm, tpd = _sort_coo(colm, node, pmv, pdv)
        self.pt_rowptr, self.pt_col, self.pt_pm, self.pt_pd = _csr_from_sorted_coo(M, tc, tn, tpm, tpd)

    # ------------------------------------------------------------------------------------
    def nbytes(self):
        return sum(getattr(self, s).nbytes for s in self.__slots__
                   if isinstance(getattr(self, s, None), np.ndarray))

    def dense(self):
        """(W, WL, Pm, Pd) for J=1 as dense numpy arrays - host bookkeeping check only."""
        def densify(n_r, n_c, rowptr, col, val):
            D = np.zeros((n_r, n_c), dtype=F32)
            r = np.repeat(np.arange(n_r), np.diff(rowptr))
            D[r, col] = val
            return D
        N, M = self.N, self.M
        W = np.zeros((N, N, 3), dtype=F32)
        W[:, :, 0] = np.eye(N, dtype=F32)
        W[:, :, 1] = np.diag(self.deg)
        W[:, :, 2] = densify(N, N, self.a_rowptr, self.a_col, self.a_val)
        if not self.dual:
            return W
        WL = np.zeros((M, M, 3), dtype=F32)
        WL[:, :, 0] = np.eye(M, dtype=F32)
        WL[:, :, 1] = np.diag(self.dl)
        WL[:, :, 2] = densify(M, M, self.b_rowptr, self.b_col, self.b_val)
        Pm = densify(N, M, self.p_rowptr, self.p_col, self.p_pm)
        Pd = densify(N, M, self.p_rowptr, self.p_col, self.p_pd)
        return W, WL, Pm, Pd


_FIELDS = {
    # name -> (rowptr, row space, col space, value arrays)
    "a": ("a_rowptr", "n", "n", ("a_col", "a_val")),
    "at": ("at_rowptr", "n", "n", ("at_col", "at_val")),
    "b": ("b_rowptr", "m", "m", ("b_col", "b_val")),
    "bt": ("bt_rowptr", "m", "m", ("bt_col", "bt_val")),
    "p": ("p_rowptr", "n", "m", ("p_col", "p_pm", "p_pd")),
    "pt": ("pt_rowptr", "m", "n", ("pt_col", "pt_pm", "pt_pd")),
    "bts": ("bts_rowptr", "m", "m", ("bts_col", "bts_val")),
    "btc": ("btc_rowptr", "m", "m", ("btc_col", "btc_val")),
}


def concat_block_diagonal(graphs, dual=True, skip=(), alloc=None, defer_offsets=False):
    """Block-diagonal batch of ``GraphOps``: global row/col indices, per-graph offsets.

    Every output array is written ONCE, straight into one contiguous staging buffer (16-byte aligned
    sub-arrays).  ``alloc(nbytes)`` supplies that buffer as a numpy uint8 array (``pack.py`` passes
    pinned memory, so the staging buffer IS the host->device copy source); default: numpy memory.
    ``skip`` lists operators to leave out (``pack.py`` skips the full ``bt`` when only the run-length
    split ``bts`` is needed).

    ``defer_offsets=False``: index arrays get their row / column / nnz offset added on the host.
    ``defer_offsets=True`` (the fast path): the per-graph arrays are copied RAW (one
    ``np.concatenate`` per field - no per-graph arithmetic on the host) and the buffer additionally
    carries a fix-up table ``fixup`` (n_entries x 4 int32: array offset, length, segment-pointer
    offset, segment-addend offset, all in 4-byte elements from the buffer start) plus the segment
    tables it refers to; ``hgnn_fixup_offsets`` applies it on the GPU after the copy
    (``apply_fixups`` is the numpy twin used by the CPU tests).

    Returns ``(arrays, buffer, layout)``: ``arrays`` = dict of numpy views (``node_off``/``edge_off``
    (bs+1), ``deg``, ``dl``, ``pad_n`` and, per operator, ``<op>_rowptr`` + index/value arrays),
    ``layout`` = {name: (byte offset, dtype, length)}."""
    bs = len(graphs)
    n = np.array([g.N for g in graphs], dtype=np.int64)
    m = np.array([g.M for g in graphs], dtype=np.int64)
    node_off = np.concatenate([[0], np.cumsum(n)])
    edge_off = np.concatenate([[0], np.cumsum(m)])
    off = {"n": node_off, "m": edge_off}
    # spec: name -> (dtype, [source arrays], drop last element of each, per-graph addend or None, tail)
    spec = {"node_off": (I32, [node_off.astype(I32)], False, None, None),
            "edge_off": (I32, [edge_off.astype(I32)], False, None, None),
            "pad_n": (F32, [(int(n.max()) - n).astype(F32) if bs else np.zeros(0, F32)], False, None, None),
            "deg": (F32, [g.deg for g in graphs], False, None, None)}
    seg = {"n": "node_off", "m": "edge_off"}            # segment-pointer arrays by name
    fix = []                                            # (array, segment pointers, addends)

    def seg_table(name, values):
        spec[name] = (I32, [np.asarray(values, dtype=I32)], False, None, None)
        return name

    names = [k for k in (list(_FIELDS) if dual else ["a", "at"]) if k not in skip]
    if dual:
        spec["dl"] = (F32, [g.dl for g in graphs], False, None, None)
        nr_off = np.concatenate([[0], np.cumsum([g.bts_rng_lo.shape[0] for g in graphs])])
        ne_off = np.concatenate([[0], np.cumsum([g.bts_rng_id.shape[0] for g in graphs])])
        seg_table("_seg_rng_entries", ne_off)
        seg_table("_seg_rng_ranges", nr_off)
        spec["bts_rng_rowptr"] = (I32, [g.bts_rng_rowptr for g in graphs], True, ne_off[:-1], int(ne_off[-1]))
        fix.append(("bts_rng_rowptr", "edge_off", "_seg_rng_entries"))
        spec["bts_rng_id"] = (I32, [g.bts_rng_id for g in graphs], False, nr_off[:-1], None)
        fix.append(("bts_rng_id", "_seg_rng_entries", "_seg_rng_ranges"))
        spec["bts_rng_val"] = (F32, [g.bts_rng_val for g in graphs], False, None, None)
        for k in ("bts_rng_lo", "bts_rng_hi"):
            spec[k] = (I32, [getattr(g, k) for g in graphs], False, edge_off[:-1], None)
            fix.append((k, "_seg_rng_ranges", "edge_off"))
    for name in names:
        rp_name, rspace, cspace, arrs = _FIELDS[name]
        nnz_off = np.concatenate([[0], np.cumsum([getattr(g, arrs[0]).shape[0] for g in graphs])])
        seg_nnz = seg_table("_seg_nnz_" + name, nnz_off)
        spec[rp_name] = (I32, [getattr(g, rp_name) for g in graphs], True, nnz_off[:-1], int(nnz_off[-1]))
        fix.append((rp_name, seg[rspace], seg_nnz))
        spec[arrs[0]] = (I32, [getattr(g, arrs[0]) for g in graphs], False, off[cspace][:-1], None)
        fix.append((arrs[0], seg_nnz, seg[cspace]))
        for a in arrs[1:]:
            spec[a] = (F32, [getattr(g, a) for g in graphs], False, None, None)
    if dual:     # collapsed line graph: per-row weights and the list of active rows
        spec["ew"] = (F32, [g.ew for g in graphs], False, None, None)
        na_off = np.concatenate([[0], np.cumsum([g.erow.shape[0] for g in graphs])])
        seg_table("_seg_erow", na_off)
        spec["erow"] = (I32, [g.erow for g in graphs], False, edge_off[:-1], None)
        fix.append(("erow", "_seg_erow", "edge_off"))
    # layout
    layout, total = {}, 0
    for key, (dt, parts, drop, add, tail) in spec.items():
        length = sum(p.shape[0] - (1 if drop else 0) for p in parts) + (1 if tail is not None else 0)
        layout[key] = (total, dt, length)
        total += (4 * length + 15) & ~15
    if defer_offsets:
        layout["fixup"] = (total, I32, 4 * len(fix))
        total += (16 * len(fix) + 15) & ~15
    buf = alloc(max(total, 16)) if alloc is not None else np.empty(max(total, 16), dtype=np.uint8)
    arrays = {}
    for key, (dt, parts, drop, add, tail) in spec.items():
        o, _, length = layout[key]
        view = buf[o:o + 4 * length].view(dt)
        body = length - (1 if tail is not None else 0)
        if defer_offsets or add is None:
            srcs = [p[:-1] for p in parts] if drop else parts
            if len(srcs) == 1:
                view[:body] = srcs[0]
            elif body:
                np.concatenate(srcs, out=view[:body])
        else:
            pos = 0
            for src, ad in zip(parts, add):
                k = src.shape[0] - (1 if drop else 0)
                if k:
                    np.add(src[:k], I32(ad), out=view[pos:pos + k], casting="unsafe")
                pos += k
        if tail is not None:
            view[body] = tail
        arrays[key] = view
    if defer_offsets:
        o, _, length = layout["fixup"]
        table = buf[o:o + 4 * length].view(I32).reshape(-1, 4)
        for i, (arr, segp, sadd) in enumerate(fix):
            n_body = layout[arr][2] - (1 if spec[arr][4] is not None else 0)
            table[i] = (layout[arr][0] // 4, n_body, layout[segp][0] // 4, layout[sadd][0] // 4)
        arrays["fixup"] = table
    return arrays, buf, layout


def apply_fixups(buf, layout, n_graphs):
    """numpy twin of ``hgnn_fixup_offsets``: arr[i] += addend[g] for i in [segptr[g], segptr[g+1])."""
    o, _, length = layout["fixup"]
    words = buf.view(I32)
    table = buf[o:o + 4 * length].view(I32).reshape(-1, 4)
    for arr_off, n_body, seg_off, add_off in table:
        segp = words[seg_off:seg_off + n_graphs + 1]
        addend = words[add_off:add_off + n_graphs]
        arr = words[arr_off:arr_off + n_body]
        for g in range(n_graphs):
            arr[segp[g]:segp[g + 1]] += addend[g]
