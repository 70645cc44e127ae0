"""Device-resident block-diagonal operator pack and the handles that stand in for the reference's
dense ``W / WL / Pm / Pd / mask`` tensors.

Reference boundary: functions/batching.py:77-185 returns dense ``W (bs,N,N,K)``, ``WL (bs,M,M,K)``,
``Pm, Pd (bs,N,M)`` and masks; scripts/train_mnb.py:56-70 only sets ``.requires_grad``, calls
``.cuda()`` and hands them back to the model.  ``OperatorHandle`` / ``MaskHandle`` support exactly
that, while the data stays CSR (SURVEY.md section 8b "Tensors crossing it").

HBM layout: all index/value arrays of a batch live in ONE contiguous device buffer (a single
host->device copy from pinned memory), 16-byte aligned sub-arrays: int32 rowptr/col, fp32 values,
block-diagonal with global row numbers.  At C2 (32 graphs, N=1000) that is ~10 MB instead of the
reference's 9.6 GB of dense WL.
"""
import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib
from ._lib import call, fptr, iptr, stream
from . import sparse_ops
from .sparse_ops import GraphOps, concat_block_diagonal  # noqa: F401


class Csr(object):
    """CSR triple (or quad, for the Pm/Pd pair) of device tensors."""
    __slots__ = ("rowptr", "col", "val", "val2", "n_rows")

    def __init__(self, rowptr, col, val, val2=None):
        self.rowptr, self.col, self.val, self.val2 = rowptr, col, val, val2
        self.n_rows = rowptr.numel() - 1

    @property
    def nnz(self):
        return self.col.numel()

    def desc(self, second=False):
        return ("csr", self.rowptr, self.col, self.val2 if second else self.val)


def _pinned_alloc(holder):
    """alloc(nbytes) -> numpy uint8 view of a pinned torch buffer (kept alive in ``holder``)."""
    def alloc(nbytes):
        holder["host"] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        return holder["host"].numpy()
    return alloc


class _StagingRing(object):
    """A few long-lived pinned staging buffers reused round-robin.  A fresh
    ``torch.empty(17 MB, pin_memory=True)`` per batch costs 3-4 ms whenever the caching host allocator
    cannot recycle a block (measured: profiles/README.md, prepare_probe) - ten times the copy itself.
    A slot is handed out again only after the CUDA event recorded behind its last copy completed."""

    def __init__(self, n_slots=4):
        self.bufs, self.events, self.next = [None] * n_slots, [None] * n_slots, 0
        self.lock = threading.Lock()

    def acquire(self, nbytes):
        with self.lock:
            i = self.next
            self.next = (i + 1) % len(self.bufs)
            ev, self.events[i] = self.events[i], None
        if ev is not None:
            ev.synchronize()
        buf = self.bufs[i]
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
            self.bufs[i] = buf
        return i, buf

    def copied(self, i):
        """Call after enqueueing the copy out of slot i on the current stream."""
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


_staging = _StagingRing()
_meta_ring = _StagingRing(8)


class _PinnedSlabs(object):
    """Bump allocator over pinned slabs for the graph blobs of a dataset (sparse_ops.blob_alloc):
    with the dataset in pinned memory, ``BatchPack.from_graphs`` needs no host-side concatenation
    - every blob is DMA'd as it is and the GPU assembles the batch (csrc/hostpack.cu).  Slabs are
    never recycled blob by blob; a slab is released when all graphs carved from it are gone.
    Beyond ``HGNN_B200_PINNED_DATASET_MB`` (default 16384) blobs fall back to pageable memory
    (the copies are then staged by the driver: slower, same result)."""
    SLAB = 32 << 20

    def __init__(self):
        self.cur, self.pos, self.total = None, 0, 0
        self.cap = int(os.environ.get("HGNN_B200_PINNED_DATASET_MB", "16384")) << 20
        self.lock = threading.Lock()

    def __call__(self, nbytes):
        if not torch.cuda.is_available() or self.total + nbytes > self.cap:
            return np.empty(nbytes, dtype=np.uint8)
        need = (nbytes + 63) & ~63
        with self.lock:
            if self.cur is None or self.pos + need > self.cur.shape[0]:
                size = max(self.SLAB, need)
                self.cur = torch.empty(size, dtype=torch.uint8, pin_memory=True).numpy()
                self.pos = 0
            out = self.cur[self.pos:self.pos + nbytes]
            self.pos += need
            self.total += need
        return out


sparse_ops.blob_alloc = _PinnedSlabs()


def device_pack(graphs, dual=True, skip_bt=False, device="cuda"):
    """Block-diagonal batch assembled ON THE GPU: one host->device DMA per graph blob, then one
    gather kernel that concatenates the fields and globalises the indices
    (``hgnn_pack_device_upload``).  Returns ``(views {name: device tensor}, buffer, nbytes copied)``."""
    global _HOST_KEYS
    lib = _lib.lib
    if _HOST_KEYS is None:
        _HOST_KEYS = [lib.hgnn_host_pack_key(k).decode() for k in range(lib.hgnn_host_pack_n_keys())]
    bs = len(graphs)
    blobs = (ctypes.c_void_p * max(bs, 1))(*[g.blob_ptr() for g in graphs])
    device_pack.last_blobs = (graphs, blobs)      # prepare_batch reuses the pointer table (a Python call per graph)
    lay = (ctypes.c_longlong * (2 * len(_HOST_KEYS)))()
    stage_b, meta_b = ctypes.c_longlong(), ctypes.c_longlong()
    du, sk = 1 if dual else 0, 1 if skip_bt else 0
    total = lib.hgnn_pack_device_plan(bs, blobs, du, sk, lay, ctypes.byref(stage_b), ctypes.byref(meta_b))
    if total < 0:
        raise RuntimeError("hgnn_pack_device_plan failed: %s" % lib.hgnn_last_error().decode())
    buf = torch.empty(total, dtype=torch.uint8, device=device)
    stage = torch.empty(stage_b.value, dtype=torch.uint8, device=device)
    meta_dev = torch.empty(meta_b.value, dtype=torch.uint8, device=device)
    slot, meta_host = _meta_ring.acquire(meta_b.value)
    _lib.call("hgnn_pack_device_upload", bs, blobs, du, sk, buf.data_ptr(), stage.data_ptr(), meta_host.data_ptr(),
              meta_dev.data_ptr(), stream())
    _meta_ring.copied(slot)
    # all array views with ONE split per dtype view (a Python slicing op per array costs ~2 us x 49 arrays)
    sizes, names, pos = [], [], 0
    for k, name in enumerate(_HOST_KEYS):
        n = lay[2 * k + 1]
        if n < 0:
            continue
        o = lay[2 * k] >> 2
        if o > pos:
            sizes.append(o - pos)
            names.append(None)
        sizes.append(n)
        names.append(name)
        pos = o + n
    tail = (total >> 2) - pos
    if tail > 0:
        sizes.append(tail)
        names.append(None)
    parts_i = buf.view(torch.int32).split_with_sizes(sizes)
    parts_f = buf.view(torch.float32).split_with_sizes(sizes)
    views = {name: (parts_f[i] if name in _FLOAT_KEYS else parts_i[i]) for i, name in enumerate(names) if name is not None}
    return views, buf, stage_b.value + meta_b.value


_HOST_KEYS = None
_FLOAT_KEYS = frozenset(["pad_n", "deg", "dl", "a_val", "at_val", "b_val", "bt_val", "bts_val", "bts_rng_val", "btc_val", "ew",
                         "p_pm", "p_pd", "pt_pm", "pt_pd"])
HOST_PACK_THREADS = max(1, min(8, (os.cpu_count() or 1) // 2))
# True (env HGNN_B200_HOST_CONCAT=1): concatenate a batch on the host (host_pack) and copy it once,
# instead of the default per-graph DMA + GPU gather (device_pack).  Same arrays either way.
HOST_CONCAT = os.environ.get("HGNN_B200_HOST_CONCAT", "0") == "1"


def host_pack(graphs, dual=True, skip_bt=False, alloc=None, n_threads=None):
    """Block-diagonal batch of ``GraphOps`` through the native packer (csrc/hostpack.cu): every
    graph's blob is concatenated field by field into ONE staging buffer by a small thread pool.
    Same arrays and fix-up table as ``sparse_ops.concat_block_diagonal(defer_offsets=True)``
    (functions/batching.py:77-185 is the reference's dense zero-padding).  Host code only - works
    without a GPU.  Returns ``(buffer (numpy uint8), layout {name: (byte offset, dtype, length)})``."""
    global _HOST_KEYS
    lib = _lib.lib
    if _HOST_KEYS is None:
        _HOST_KEYS = [lib.hgnn_host_pack_key(k).decode() for k in range(lib.hgnn_host_pack_n_keys())]
    bs = len(graphs)
    blobs = (ctypes.c_void_p * max(bs, 1))(*[g.blob_ptr() for g in graphs])
    lay = (ctypes.c_longlong * (2 * len(_HOST_KEYS)))()
    total = lib.hgnn_host_pack_layout(bs, blobs, 1 if dual else 0, 1 if skip_bt else 0, lay)
    if total < 0:
        raise RuntimeError("hgnn_host_pack_layout failed: %s" % lib.hgnn_last_error().decode())
    buf = alloc(total) if alloc is not None else np.empty(total, dtype=np.uint8)
    rc = lib.hgnn_host_pack_fill(bs, blobs, 1 if dual else 0, 1 if skip_bt else 0, lay, buf.ctypes.data,
                                 HOST_PACK_THREADS if n_threads is None else int(n_threads))
    if rc != 0:
        raise RuntimeError("hgnn_host_pack_fill failed: %s" % lib.hgnn_last_error().decode())
    layout = {}
    for k, name in enumerate(_HOST_KEYS):
        if lay[2 * k + 1] >= 0:
            layout[name] = (lay[2 * k], np.float32 if name in _FLOAT_KEYS else np.int32, lay[2 * k + 1])
    return buf, layout


def _device_views(host_tensor, layout, device):
    """One H2D copy of the staging buffer; the arrays become views of the device buffer."""
    dev = torch.empty(host_tensor.numel(), dtype=torch.uint8, device=device)
    dev.copy_(host_tensor, non_blocking=True)
    out = {}
    for k, (o, dt, length) in layout.items():
        tdt = torch.int32 if dt is np.int32 else torch.float32
        out[k] = dev[o:o + 4 * length].view(tdt)
    return out, dev


def exclusive_scan(counts):
    """counts (n,) int32 device -> (n+1,) int32 exclusive prefix sums (hgnn_exclusive_scan_i32)."""
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int32, device=counts.device)
    call("hgnn_exclusive_scan_i32", iptr(counts), iptr(out), n, stream())
    return out


def spgemm(A, B, clip=False):
    """C = A @ B for block-diagonal CSR operands (functions/operators.py:26-29, 78-81)."""
    R = A.n_rows
    dev = A.rowptr.device
    if R == 0 or A.nnz == 0 or B.nnz == 0:
        return Csr(torch.zeros(R + 1, dtype=torch.int32, device=dev),
                   torch.empty(0, dtype=torch.int32, device=dev),
                   torch.empty(0, dtype=torch.float32, device=dev))
    cnt = torch.empty(R, dtype=torch.int32, device=dev)
    call("hgnn_spgemm_count_products", R, iptr(A.rowptr), iptr(A.col), iptr(B.rowptr), iptr(cnt), stream())
    prodptr = exclusive_scan(cnt)
    total = int(prodptr[-1].item())
    pcol = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    pval = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    pflag = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    rowcnt = torch.empty(R, dtype=torch.int32, device=dev)
    call("hgnn_spgemm_expand", R, iptr(A.rowptr), iptr(A.col), fptr(A.val), iptr(B.rowptr), iptr(B.col),
         fptr(B.val), iptr(prodptr), iptr(pcol), fptr(pval), iptr(pflag), iptr(rowcnt), stream())
    c_rowptr = exclusive_scan(rowcnt)
    nnz = int(c_rowptr[-1].item())
    c_col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    c_val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz]
    call("hgnn_spgemm_fill", R, iptr(prodptr), iptr(pcol), fptr(pval), iptr(pflag), iptr(c_rowptr),
         iptr(c_col), fptr(c_val), 1 if clip else 0, stream())
    return Csr(c_rowptr, c_col, c_val)


def dense_to_csr(D1, D2, bs, row_off, col_off, n_rows):
    """Strided dense (bs, rows, cols) view(s) -> CSR on the device.  Pattern = nz(D1) | nz(D2)."""
    dev = D1.device
    sb, sr, sc = D1.stride()
    if D2 is not None and D2.stride() != D1.stride():
        raise RuntimeError("hgnn_b200: Pm and Pd must share one memory layout")
    if n_rows == 0 or D1.numel() == 0:
        z = torch.zeros(n_rows + 1, dtype=torch.int32, device=dev)
        e = torch.empty(0, dtype=torch.float32, device=dev)
        return Csr(z, torch.empty(0, dtype=torch.int32, device=dev), e, e if D2 is not None else None)
    cnt = torch.empty(n_rows, dtype=torch.int32, device=dev)
    p1 = D1.data_ptr()
    p2 = D2.data_ptr() if D2 is not None else None
    call("hgnn_dense_count_nnz", p1, p2, sb, sr, sc, bs, iptr(row_off), iptr(col_off), iptr(cnt), stream())
    rowptr = exclusive_scan(cnt)
    nnz = int(rowptr[-1].item())
    col = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    v1 = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz]
    v2 = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)[:nnz] if D2 is not None else None
    call("hgnn_dense_fill_csr", p1, p2, sb, sr, sc, bs, iptr(row_off), iptr(col_off), iptr(rowptr),
         iptr(col), fptr(v1), fptr(v2), stream())
    return Csr(rowptr, col, v1, v2)


class BatchPack(object):
    """Block-diagonal CSR operators of one batch, resident in HBM.

    ``node_ops()`` / ``edge_ops()`` give the K = J+2 operator descriptors in the reference's channel
    order [I, D, A, A^2, ...] (layers_mnb.py:409); ``*_T`` the transposed operators the backward
    kernels gather with.  ``p`` (rows = nodes) and ``pt`` (rows = line-graph nodes) hold Pm and Pd
    on one sparsity pattern.
    """

    def __init__(self):
        self.J = 1
        self.dual = True
        self.generic = False     # True: built from dense tensors, all K operators are plain CSR
        self.requires_grad = False

    # ---- construction from host GraphOps (the fast path) ----------------------------------
    @classmethod
    def from_graphs(cls, graphs, J=1, dual=True, device="cuda", clip_powers=False):
        _lib.require_cuda()
        self = cls()
        device = torch.device(device)
        self.J, self.dual, self.device = int(J), bool(dual), device
        # The full transposed line-graph operator is only needed by the layer-level kernels and by
        # the powers; the engine uses its run-length split twin.  It is uploaded on first use.
        lazy_bt = dual and self.J == 1
        if HOST_CONCAT:
            # raw per-graph arrays -> pinned staging slot (native multi-threaded concat of the graph
            # blobs); the per-graph index offsets are added on the GPU after the copy
            slot = {}

            def staging(nbytes):
                slot["i"], slot["buf"] = _staging.acquire(nbytes)
                slot["n"] = nbytes
                return slot["buf"].numpy()[:nbytes]

            _, layout = host_pack(graphs, dual=dual, skip_bt=lazy_bt, alloc=staging)
        self.bs = len(graphs)
        self.n_nodes = np.array([g.N for g in graphs], dtype=np.int64)
        self.n_edges = np.array([g.M for g in graphs], dtype=np.int64)
        self.Nmax = int(self.n_nodes.max())
        self.Emax = int(self.n_edges.max()) if dual else 0
        self.Rn, self.Rm = int(self.n_nodes.sum()), int(self.n_edges.sum())
        if HOST_CONCAT:
            dev, self._buffer = _device_views(slot["buf"][:slot["n"]], layout, device)
            _staging.copied(slot["i"])
            self.nbytes = slot["n"]
            call("hgnn_fixup_offsets", self._buffer.data_ptr(), iptr(dev["fixup"]), dev["fixup"].numel() // 4,
                 len(graphs), stream())
        else:
            # the product path: blobs DMA'd as they are, batch assembled by one gather kernel
            dev, self._buffer, self.nbytes = device_pack(graphs, dual=dual, skip_bt=lazy_bt, device=device)
        self._host_graphs = graphs       # keeps the blobs alive behind the asynchronous copies
        self.node_off, self.edge_off = dev["node_off"], dev["edge_off"]
        self.pad_n = dev["pad_n"]
        self.deg = dev["deg"]
        self.a = [Csr(dev["a_rowptr"], dev["a_col"], dev["a_val"])]
        self.at = [Csr(dev["at_rowptr"], dev["at_col"], dev["at_val"])]
        if dual:
            self.dl = dev["dl"]
            self.b = [Csr(dev["b_rowptr"], dev["b_col"], dev["b_val"])]
            self._bt = None if lazy_bt else [Csr(dev["bt_rowptr"], dev["bt_col"], dev["bt_val"])]
            self.p = Csr(dev["p_rowptr"], dev["p_col"], dev["p_pm"], dev["p_pd"])
            self.pt = Csr(dev["pt_rowptr"], dev["pt_col"], dev["pt_pm"], dev["pt_pd"])
            # run-length split twin of bt for the engine kernels (phantom ranges stored once)
            self.bts = Csr(dev["bts_rowptr"], dev["bts_col"], dev["bts_val"])
            self.bts_ranges = (dev["bts_rng_rowptr"], dev["bts_rng_id"], dev["bts_rng_val"],
                               dev["bts_rng_lo"], dev["bts_rng_hi"])
            # collapsed line graph for the persistent engine kernels (sparse_ops.GraphOps._build_collapsed):
            # transposed operator over the active rows, per-row weights, list of active rows
            self.btc = Csr(dev["btc_rowptr"], dev["btc_col"], dev["btc_val"])
            self.ew, self.erow = dev["ew"], dev["erow"]
        for _ in range(1, self.J):   # A^(2^j): repeated squaring on the GPU, unclipped by default
            self.a.append(spgemm(self.a[-1], self.a[-1], clip_powers))
            self.at.append(spgemm(self.at[-1], self.at[-1], clip_powers))
            if dual:
                self.b.append(spgemm(self.b[-1], self.b[-1], clip_powers))
                self.bt.append(spgemm(self.bt[-1], self.bt[-1], clip_powers))
        return self

    @property
    def bt(self):
        """Transposed line-graph operator(s) as full CSR; uploaded lazily when J == 1."""
        if self._bt is None:
            gs = self._host_graphs
            nnz = np.concatenate([[0], np.cumsum([g.bt_col.shape[0] for g in gs])])
            eoff = np.concatenate([[0], np.cumsum([g.M for g in gs])])
            rp = np.concatenate([g.bt_rowptr[:-1].astype(np.int64) + nnz[i] for i, g in enumerate(gs)] + [nnz[-1:]])
            col = np.concatenate([g.bt_col.astype(np.int64) + eoff[i] for i, g in enumerate(gs)])
            val = np.concatenate([g.bt_val for g in gs])
            d = self.device
            self._bt = [Csr(torch.from_numpy(rp.astype(np.int32)).to(d), torch.from_numpy(col.astype(np.int32)).to(d),
                            torch.from_numpy(val.astype(np.float32)).to(d))]
        return self._bt

    # ---- construction from the reference's dense tensors (compatibility path) --------------
    @classmethod
    def from_dense(cls, W, WL=None, Pm=None, Pd=None, N_batch=None, E_batch=None):
        """W (bs,N,N,K) [, WL (bs,M,M,K), Pm, Pd (bs,N,M)] CUDA tensors -> generic CSR operators.
        Every one of the K slices is extracted as it is (no assumption that slice 0 is I)."""
        _lib.require_cuda()
        self = cls()
        self.generic = True
        self.device = W.device
        self.dual = WL is not None
        self.bs, self.Nmax, _, K = W.shape
        self.J = K - 2
        n = N_batch.detach().to("cpu", torch.int64).numpy()
        self.n_nodes = n
        self.Rn = int(n.sum())
        off = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
        self.node_off = torch.from_numpy(off).to(self.device)
        self.pad_n = torch.from_numpy((self.Nmax - n).astype(np.float32)).to(self.device)
        self.gen_node = [dense_to_csr(W[:, :, :, k], None, self.bs, self.node_off, self.node_off, self.Rn)
                         for k in range(K)]
        self.gen_node_T = [dense_to_csr(W[:, :, :, k].transpose(1, 2), None, self.bs, self.node_off,
                                        self.node_off, self.Rn) for k in range(K)]
        if self.dual:
            e = E_batch.detach().to("cpu", torch.int64).numpy()
            self.n_edges = e
            self.Emax = WL.shape[1]
            self.Rm = int(e.sum())
            eoff = np.concatenate([[0], np.cumsum(e)]).astype(np.int32)
            self.edge_off = torch.from_numpy(eoff).to(self.device)
            self.gen_edge = [dense_to_csr(WL[:, :, :, k], None, self.bs, self.edge_off, self.edge_off,
                                          self.Rm) for k in range(K)]
            self.gen_edge_T = [dense_to_csr(WL[:, :, :, k].transpose(1, 2), None, self.bs, self.edge_off,
                                            self.edge_off, self.Rm) for k in range(K)]
            self.p = dense_to_csr(Pm, Pd, self.bs, self.node_off, self.edge_off, self.Rn)
            self.pt = dense_to_csr(Pm.transpose(1, 2), Pd.transpose(1, 2), self.bs, self.edge_off,
                                   self.node_off, self.Rm)
        else:
            self.n_edges, self.Emax, self.Rm = np.zeros(self.bs, np.int64), 0, 0
        return self

    def record_stream(self, stream):
        """Tie EVERY device allocation owned by this pack to ``stream`` (the consumer's), for packs built on
        another stream (BatchLoader's copy stream): the main buffer, and the separately allocated SpGEMM
        powers A^(2^j) / B^(2^j) and their transposes when J > 1.  Without it a dropped batch returns those
        blocks to the producer stream's pool while consumer kernels may still be reading them."""
        seen = set()

        def visit(o, depth=0):
            if torch.is_tensor(o):
                if o.is_cuda and o.untyped_storage().data_ptr() not in seen:
                    seen.add(o.untyped_storage().data_ptr())
                    o.record_stream(stream)
            elif isinstance(o, Csr):
                for k in Csr.__slots__:
                    visit(getattr(o, k, None), depth + 1)
            elif isinstance(o, (list, tuple)) and depth < 4:
                for v in o:
                    visit(v, depth + 1)
            elif isinstance(o, dict) and depth < 4:
                for v in o.values():
                    visit(v, depth + 1)

        buf = self.__dict__.get("_buffer")
        if buf is not None and torch.is_tensor(buf) and buf.is_cuda:
            # every array of the batch assembled by device_pack is a view of this ONE buffer: register it once and walk only
            # the allocations made after it - the SpGEMM powers (J > 1), the lazily uploaded transposed operator, caches -
            # instead of asking ~50 views for their storage (75 us per batch on the consumer thread)
            seen.add(buf.untyped_storage().data_ptr())
            buf.record_stream(stream)
            own = ("node_off", "edge_off", "pad_n", "deg", "dl", "p", "pt", "bts", "bts_ranges", "btc", "ew", "erow", "_buffer")
            for key, v in self.__dict__.items():
                if key in own or key == "_host_graphs":
                    continue
                if key in ("a", "at", "b"):
                    visit(v[1:])
                else:
                    visit(v)
            return
        for key, v in self.__dict__.items():
            if key != "_host_graphs":
                visit(v)

    # ---- operator descriptor lists ------------------------------------------------------------
    def node_ops(self):
        if self.generic:
            return [c.desc() for c in self.gen_node]
        return [("ident",), ("diag", self.deg)] + [c.desc() for c in self.a]

    def node_ops_T(self):
        if self.generic:
            return [c.desc() for c in self.gen_node_T]
        return [("ident",), ("diag", self.deg)] + [c.desc() for c in self.at]

    def edge_ops(self):
        if self.generic:
            return [c.desc() for c in self.gen_edge]
        return [("ident",), ("diag", self.dl)] + [c.desc() for c in self.b]

    def edge_ops_T(self, split=False):
        """``split=True`` (engine kernels): the first-power operator as direct CSR + range entries."""
        if self.generic:
            return [c.desc() for c in self.gen_edge_T]
        if split and getattr(self, "bts", None) is not None:
            rest = self._bt[1:] if self._bt is not None else []      # J == 1: the full bt stays on the host
            return [("ident",), ("diag", self.dl), self.bts.desc() + (self.bts_ranges,)] + [c.desc() for c in rest]
        return [("ident",), ("diag", self.dl)] + [c.desc() for c in self.bt]

    @property
    def K(self):
        return self.J + 2

    # ---- densify (bit-exact operator checks, sparse=False batches) ------------------------------
    def _scatter(self, csr, second, D, row_off, col_off):
        if csr.n_rows == 0 or csr.nnz == 0 or D.numel() == 0:
            return
        sb, sr, sc = D.stride()
        call("hgnn_csr_to_dense", iptr(csr.rowptr), iptr(csr.col), fptr(csr.val2 if second else csr.val),
             self.bs, iptr(row_off), iptr(col_off), D.data_ptr(), sb, sr, sc, stream())

    def dense_W(self):
        W = torch.zeros(self.bs, self.Nmax, self.Nmax, self.K, device=self.device)
        if self.generic:
            for k, c in enumerate(self.gen_node):
                self._scatter(c, False, W[:, :, :, k], self.node_off, self.node_off)
            return W
        eye = _diag_csr(self.Rn, None, self.device)
        self._scatter(eye, False, W[:, :, :, 0], self.node_off, self.node_off)
        self._scatter(_diag_csr(self.Rn, self.deg, self.device), False, W[:, :, :, 1], self.node_off, self.node_off)
        for j, c in enumerate(self.a):
            self._scatter(c, False, W[:, :, :, 2 + j], self.node_off, self.node_off)
        return W

    def dense_WL(self):
        WL = torch.zeros(self.bs, self.Emax, self.Emax, self.K, device=self.device)
        if self.generic:
            for k, c in enumerate(self.gen_edge):
                self._scatter(c, False, WL[:, :, :, k], self.edge_off, self.edge_off)
            return WL
        self._scatter(_diag_csr(self.Rm, None, self.device), False, WL[:, :, :, 0], self.edge_off, self.edge_off)
        self._scatter(_diag_csr(self.Rm, self.dl, self.device), False, WL[:, :, :, 1], self.edge_off, self.edge_off)
        for j, c in enumerate(self.b):
            self._scatter(c, False, WL[:, :, :, 2 + j], self.edge_off, self.edge_off)
        return WL

    def dense_P(self, second):
        P = torch.zeros(self.bs, self.Nmax, self.Emax, device=self.device)
        self._scatter(self.p, second, P, self.node_off, self.edge_off)
        return P

    def dense_mask(self, lg=False):
        n = torch.as_tensor(self.n_edges if lg else self.n_nodes, device=self.device)
        size = self.Emax if lg else self.Nmax
        keep = (torch.arange(size, device=self.device).view(1, -1) < n.view(-1, 1)).float()
        return keep.unsqueeze(2) * keep.unsqueeze(1)


def _diag_csr(R, vec, device):
    idx = torch.arange(R + 1, dtype=torch.int32, device=device)
    val = torch.ones(R, device=device) if vec is None else vec
    return Csr(idx, idx[:R].contiguous(), val.contiguous())


# --------------------------------------------------------------------------------------------
# handles that quack like the reference's dense tensors
# --------------------------------------------------------------------------------------------


class OperatorHandle(object):
    """Stands in for one of W / WL / Pm / Pd.  ``name`` in {"W","WL","Pm","Pd"}; ``transposed``
    mirrors ``Pm.transpose(2, 1)`` (layers_mnb.py:214-215)."""

    def __init__(self, pack, name, transposed=False):
        self.pack, self.name, self.transposed = pack, name, transposed
        self.requires_grad = False

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    @property
    def is_cuda(self):
        return True

    @property
    def device(self):
        return self.pack.device

    @property
    def shape(self):
        p = self.pack
        if self.name == "W":
            return torch.Size([p.bs, p.Nmax, p.Nmax, p.K])
        if self.name == "WL":
            return torch.Size([p.bs, p.Emax, p.Emax, p.K])
        return torch.Size([p.bs, p.Emax, p.Nmax] if self.transposed else [p.bs, p.Nmax, p.Emax])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def transpose(self, d0, d1):
        if self.name not in ("Pm", "Pd") or {d0 % 3, d1 % 3} != {1, 2}:
            raise RuntimeError("hgnn_b200: only Pm/Pd.transpose(2, 1) is defined on operator handles")
        return OperatorHandle(self.pack, self.name, not self.transposed)

    def to_dense(self):
        """The reference's dense tensor, produced on the GPU from the CSR pack (bit-exact)."""
        if self.name == "W":
            return self.pack.dense_W()
        if self.name == "WL":
            return self.pack.dense_WL()
        P = self.pack.dense_P(self.name == "Pd")
        return P.transpose(2, 1) if self.transposed else P

    def __repr__(self):
        return "OperatorHandle(%s%s, shape=%s)" % (self.name, ".T" if self.transposed else "", tuple(self.shape))


class MaskHandle(object):
    """Stands in for ``mask`` / ``mask_lg`` (batching.py:182-183) without the (bs, M, M) tensor."""

    def __init__(self, pack, lg):
        self.pack, self.lg = pack, lg
        self.requires_grad = False

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    @property
    def shape(self):
        s = self.pack.Emax if self.lg else self.pack.Nmax
        return torch.Size([self.pack.bs, s, s])

    def to_dense(self):
        return self.pack.dense_mask(self.lg)

    def __getitem__(self, idx):
        return self.to_dense()[idx]


def is_handle(x):
    return isinstance(x, (OperatorHandle, MaskHandle))


_DENSE_CACHE = {}


class PackTensor(torch.Tensor):
    """The edge feature ``XL`` as ``prepare_batch`` builds it (line-graph degree, functions/batching.py:171): an
    ordinary tensor that remembers WHICH pack's degree it holds.  The tag survives the copies a train loop makes
    (``.cuda()``, ``.to()``, ``.pin_memory()``, ``.contiguous()``, ``.detach()``, ``.clone()``; scripts/train_mnb.py:60)
    and is dropped by every other operation and by in-place writes (version counter).  The model reads it to know,
    without touching the data, that the reference's phantom line-graph rows carry identical features - the condition
    under which the engine computes one representative per graph (sparse_ops.GraphOps._build_collapsed)."""
    _KEEP = frozenset(["cuda", "to", "pin_memory", "contiguous", "detach", "clone", "float", "requires_grad_"])

    @staticmethod
    def wrap(t, pack):
        r = t.as_subclass(PackTensor)
        r._hgnn_pack, r._hgnn_version = pack, r._version
        return r

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **kwargs)
        src = args[0] if args else None
        if (getattr(func, "__name__", "") in cls._KEEP and isinstance(src, PackTensor) and torch.is_tensor(out)
                and out.dtype == torch.float32 and out.shape == src.shape and PackTensor.pack_of(src) is not None):
            if out is src:
                return out
            out = out.as_subclass(PackTensor)
            out._hgnn_pack, out._hgnn_version = src._hgnn_pack, out._version
        elif isinstance(out, PackTensor):
            out = out.as_subclass(torch.Tensor)
        return out

    @staticmethod
    def pack_of(t):
        """The pack whose line-graph degree ``t`` provably holds, else None."""
        if not isinstance(t, PackTensor):
            return None
        pack = t.__dict__.get("_hgnn_pack")
        if pack is None or t.__dict__.get("_hgnn_version") != t._version:
            return None
        return pack


def resolve_pack(W, WL=None, Pm=None, Pd=None, N_batch=None, E_batch=None):
    """The BatchPack behind a model input: handles carry it; dense tensors are converted on the
    device (cached on the identity/version of the tensors, so a loop over one batch converts once)."""
    if isinstance(W, OperatorHandle):
        return W.pack
    if not torch.is_tensor(W):
        raise TypeError("hgnn_b200: W must be a tensor or an OperatorHandle, got %r" % type(W))
    if not W.is_cuda:
        raise RuntimeError("hgnn_b200: operators must be CUDA tensors or OperatorHandles - there is "
                           "no CPU fallback (call .cuda() as scripts/train_mnb.py:60 does)")
    if N_batch is None:
        raise RuntimeError("hgnn_b200: N_batch is required to convert dense operators")
    ts = [t for t in (W, WL, Pm, Pd, N_batch, E_batch) if t is not None]
    key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts)
    hit = _DENSE_CACHE.get("last")
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        pack = BatchPack.from_dense(W.detach(), None if WL is None else WL.detach(),
                                    None if Pm is None else Pm.detach(),
                                    None if Pd is None else Pd.detach(), N_batch, E_batch)
    _DENSE_CACHE["last"] = (key, pack, ts)   # keep the tensors alive so data_ptr stays unique
    return pack


# --------------------------------------------------------------------------------------------
# per-graph handles: the sparse form of one instance's [W, WL, Pm, Pd]  
# --------------------------------------------------------------------------------------------


class SparseAdj(object):
    """Edge-list adjacency standing in for the dense ``A (N,N)`` of an instance
    (functions/data_generator.py:85 format ``[X, A, y, W, WL, Pm, Pd]``) when N is too large for a
    dense matrix.  ``rows, cols, vals`` list every stored entry (both directions of each edge)."""

    def __init__(self, N, rows, cols, vals):
        self.N = int(N)
        self.rows = np.asarray(rows, dtype=np.int64)
        self.cols = np.asarray(cols, dtype=np.int64)
        self.vals = np.asarray(vals, dtype=np.float32)

    @property
    def shape(self):
        return torch.Size([self.N, self.N])

    def nnz(self):
        return int(np.count_nonzero(self.vals))

    def to_dense(self):
        A = torch.zeros(self.N, self.N)
        A[torch.from_numpy(self.rows), torch.from_numpy(self.cols)] = torch.from_numpy(self.vals)
        return A

    def __add__(self, other):     # scripts/train_ccn.py:36 does ``A + torch.eye(N)``
        return self.to_dense() + other


class GraphHandle(object):
    """One of W / WL / Pm / Pd of a single graph, backed by host ``GraphOps`` (sparse instance)."""

    def __init__(self, graph_ops, name, J=1):
        self.graph_ops, self.name, self.J = graph_ops, name, J
        self.requires_grad = False

    @property
    def shape(self):
        g = self.graph_ops
        return {"W": torch.Size([g.N, g.N, self.J + 2]), "WL": torch.Size([g.M, g.M, self.J + 2]),
                "Pm": torch.Size([g.N, g.M]), "Pd": torch.Size([g.N, g.M])}[self.name]

    def to_dense(self):
        """Dense tensor (CPU) identical to the reference's graph_operators output."""
        pack = BatchPack.from_graphs([self.graph_ops], self.J, dual=self.graph_ops.dual)
        h = OperatorHandle(pack, self.name)
        return h.to_dense()[0].cpu()

    def __repr__(self):
        return "GraphHandle(%s, shape=%s)" % (self.name, tuple(self.shape))
