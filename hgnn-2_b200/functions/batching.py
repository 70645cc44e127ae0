"""Mini-batch assembly - mirror of the reference's functions/batching.py (:52-74, :77-185).

``prepare_batch`` keeps the reference's signature and 11-tuple return order
``X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch`` (:185).  Instead of zero-padding dense
operator tensors (9.6 GB of WL for 32 graphs of N=1000) the per-graph CSR operators are
concatenated block-diagonally and copied to the GPU once; the ``W/WL/Pm/Pd/mask/mask_lg`` slots hold
light handles that accept ``.requires_grad = ...`` and ``.cuda()`` exactly as
scripts/train_mnb.py:56-66 uses them.  ``sparse=False`` returns the reference's dense tensors
(built on the GPU from the same pack, bit-exact).
"""
import ctypes
import os
import weakref
from random import shuffle

import numpy as np
import torch

from .. import _lib
from ..pack import BatchPack, GraphHandle, MaskHandle, OperatorHandle, PackTensor, device_pack
from .operators import graph_ops_of


def _divide_batch(nb_samples_in, batch_size, idx):
    """Consecutive slices of ``idx``; the last batch takes the remainder (reference :26-40)."""
    nb_batches = -(-nb_samples_in // batch_size)
    return [idx[i * batch_size:] if i == nb_batches - 1 else idx[i * batch_size:(i + 1) * batch_size]
            for i in range(nb_batches)]


def get_batches(nb_samples_in, batch_size, data, shuffle_batch=False, sort_batch=False):
    """Batch index lists (reference :52-74): unsorted (optionally shuffled samples), or sorted by
    node count with optionally shuffled batch order."""
    if sort_batch == False:  # noqa: E712  (the reference compares with ==)
        idx = list(range(nb_samples_in))
        if shuffle_batch == True:  # noqa: E712
            shuffle(idx)
        return _divide_batch(nb_samples_in, batch_size, idx)
    sizes = np.array([data[i][0].shape[0] for i in range(len(data))], dtype=np.float64)
    idx_list = _divide_batch(nb_samples_in, batch_size, np.argsort(sizes))
    if shuffle_batch == True:  # noqa: E712
        shuffle(idx_list)
    return idx_list


_OPS_CACHE = {}


def _instance_ops(inst):
    """Host GraphOps of one instance ``[x, A, t, W, WL, Pm, Pd]``: carried by a sparse instance's
    handles, otherwise derived from A once and cached on the identity of the A tensor."""
    w = inst[3] if len(inst) > 3 else None
    if isinstance(w, GraphHandle) and w.graph_ops.dual:
        return w.graph_ops
    A = inst[1]
    key = id(A)
    hit = _OPS_CACHE.get(key)
    if hit is not None and hit[0]() is A:
        return hit[1]
    g = graph_ops_of(A, dual=True)
    try:
        _OPS_CACHE[key] = (weakref.ref(A, lambda _r, k=key: _OPS_CACHE.pop(k, None)), g)
    except TypeError:
        pass
    return g


def prepare_batch(batch, task, J=1, sparse=None, device="cuda"):
    """batch: list of ``[x (N,F), A, t, W, WL, Pm, Pd]`` instances (reference :77-95).

    Returns the reference's 11-tuple.  ``X (bs,F,Nmax)``, ``T (bs,1)``, ``XL (bs,1,Emax)``
    (= line-graph degree, :171), ``N_batch``, ``E_batch`` are CPU tensors like the reference's; the
    operator and mask slots are handles over one device-resident ``BatchPack`` unless
    ``sparse=False`` (or env HGNN_B200_DENSE_BATCH=1).
    """
    if sparse is None:
        sparse = os.environ.get("HGNN_B200_DENSE_BATCH", "0") != "1"
    bs = len(batch)
    graphs = [_instance_ops(inst) for inst in batch]
    # the operator pack first: its host->device DMAs (one per graph blob) then run while the rest is assembled
    pack = BatchPack.from_graphs(graphs, J, dual=True, device=device)
    n_feat = batch[0][0].shape[1]
    N_batch = torch.tensor([g.N for g in graphs], dtype=torch.int64)
    E_batch = torch.tensor([g.M for g in graphs], dtype=torch.int64)
    Nmax, Emax = int(N_batch.max()), int(E_batch.max())
    # host tensors in pinned memory (when CUDA is there) so that the caller's .cuda() is one DMA each
    pin = torch.cuda.is_available()
    # uninitialised + explicit zero tails: a pinned torch.zeros of the two padded tensors (1.3 MB on C2) costs ~0.2 ms of
    # host memset per batch although almost every slot is overwritten below
    X = torch.empty(bs, n_feat, Nmax, pin_memory=pin)
    XL = torch.empty(bs, 1, Emax, pin_memory=pin)
    ts = [inst[2] for inst in batch]
    if all(torch.is_tensor(t) and t.dim() == 1 and t.shape == ts[0].shape and t.dtype == ts[0].dtype for t in ts):
        T = torch.stack(ts, 0)[:, task].to(torch.float32).reshape(bs, 1)      # one gather instead of bs item() calls
    else:
        T = torch.tensor([float(inst[2][task]) for inst in batch], dtype=torch.float32).view(bs, 1)
    xs = [inst[0] for inst in batch]
    f32, row_major = torch.float32, (n_feat, 1)
    if all(type(x) is torch.Tensor and x.dtype is f32 and x.stride() == row_major and x.shape[1] == n_feat
           and not x.is_cuda for x in xs):
        # one foreign call: transposed copy of every graph's (N, F) features, line-graph degrees straight from the graph
        # blobs, zero tails (csrc/hostpack.cu: hgnn_host_fill_features)
        last = getattr(device_pack, "last_blobs", None)
        if last is not None and last[0] is graphs:
            blobs = last[1]
        else:
            blobs = (ctypes.c_void_p * bs)(*[g.blob_ptr() for g in graphs])
        rows = (ctypes.c_void_p * bs)(*[x.data_ptr() for x in xs])
        rc = _lib.lib.hgnn_host_fill_features(bs, blobs, rows, n_feat, Nmax, X.data_ptr(), Emax, XL.data_ptr())
        if rc != 0:
            raise RuntimeError("hgnn_host_fill_features failed (%d): %s" % (rc, _lib.lib.hgnn_last_error().decode()))
    else:
        Xn, XLn = X.numpy(), XL.numpy()
        for i, inst in enumerate(batch):
            g = graphs[i]
            Xn[i, :, :g.N] = np.asarray(inst[0], dtype=np.float32).T
            XLn[i, 0, :g.M] = g.dl
            if g.N < Nmax:
                Xn[i, :, g.N:] = 0.0
            if g.M < Emax:
                XLn[i, 0, g.M:] = 0.0
    XL = PackTensor.wrap(XL, pack)      # remembers that it is this pack's line-graph degree (pack.PackTensor)
    W, WL = OperatorHandle(pack, "W"), OperatorHandle(pack, "WL")
    Pm, Pd = OperatorHandle(pack, "Pm"), OperatorHandle(pack, "Pd")
    mask, mask_lg = MaskHandle(pack, False), MaskHandle(pack, True)
    if not sparse:
        W, WL, Pm, Pd = (h.to_dense().cpu() for h in (W, WL, Pm, Pd))
        mask, mask_lg = mask.to_dense().cpu(), mask_lg.to_dense().cpu()
    return X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch


class BatchLoader(object):
    """Prepared batches with one batch of look-ahead: ``prepare_batch`` (host concat of the graph
    blobs + the pinned host->device copy + the GPU offset fix-up) for batch k+1 runs on a background
    thread and a dedicated copy stream while the caller trains on batch k.

        for X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch in BatchLoader(data, idx_lists, task, J):
            ...   # exactly the body of the reference's loop (scripts/train_mnb.py:43-70)

    ``index_lists`` is what ``get_batches`` returns.  Each yielded tuple is what
    ``prepare_batch([data[i] for i in idx], task, J)`` returns; the consumer's current stream is made
    to wait for the copy, and the pack buffer is registered with it (``record_stream``), so no
    synchronisation is needed in user code."""

    def __init__(self, data, index_lists, task, J=1, depth=2, device="cuda"):
        self.data, self.index_lists, self.task, self.J = data, list(index_lists), task, J
        self.depth, self.device = max(1, int(depth)), torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

    def __len__(self):
        return len(self.index_lists)

    def __iter__(self):
        import queue
        import threading
        q = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        copy_stream = torch.cuda.Stream(device=self.device)

        def produce():
            try:
                torch.cuda.set_device(self.device)
                for idx in self.index_lists:
                    if stop.is_set():
                        return
                    with torch.cuda.stream(copy_stream):
                        batch = prepare_batch([self.data[i] for i in idx], self.task, self.J, device=self.device)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    q.put((batch, ev))
                q.put(None)
            except BaseException as e:      # surfaced in the consumer thread
                q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                batch, ev = item
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                batch[1].pack.record_stream(cur)          # every allocation of the pack, incl. J > 1 powers
                for t_ in batch:
                    if torch.is_tensor(t_) and t_.is_cuda:
                        t_.record_stream(cur)
                yield batch
        finally:
            stop.set()
            while t.is_alive():             # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                t.join(timeout=0.01)
