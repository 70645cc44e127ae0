"""Mini-batch assembly - mirror of the reference's functions/batching.py (:52-74, :77-185).

``prepare_batch`` keeps the reference's signature and 11-tuple return order
``X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch`` (:185).  Instead of zero-padding dense
operator tensors (9.6 GB of WL for 32 graphs of N=1000) the per-graph CSR operators are
concatenated block-diagonally and copied to the GPU once; the ``W/WL/Pm/Pd/mask/mask_lg`` slots hold
light handles that accept ``.requires_grad = ...`` and ``.cuda()`` exactly as
scripts/train_mnb.py:56-66 uses them.  ``sparse=False`` returns the reference's dense tensors
(built on the GPU from the same pack, bit-exact).
"""
import os
import weakref
from random import shuffle

import numpy as np
import torch

from ..pack import BatchPack, GraphHandle, MaskHandle, OperatorHandle
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
    n_feat = batch[0][0].shape[1]
    N_batch = torch.tensor([g.N for g in graphs], dtype=torch.int64)
    E_batch = torch.tensor([g.M for g in graphs], dtype=torch.int64)
    Nmax, Emax = int(N_batch.max()), int(E_batch.max())
    # host tensors in pinned memory (when CUDA is there) so that the caller's .cuda() is one DMA each
    pin = torch.cuda.is_available()
    X = torch.zeros(bs, n_feat, Nmax, pin_memory=pin)
    XL = torch.zeros(bs, 1, Emax, pin_memory=pin)
    T = torch.zeros(bs, 1)
    Xn, XLn = X.numpy(), XL.numpy()
    for i, inst in enumerate(batch):
        g = graphs[i]
        Xn[i, :, :g.N] = inst[0].numpy().T
        XLn[i, 0, :g.M] = g.dl
        T[i, 0] = float(inst[2][task])
    pack = BatchPack.from_graphs(graphs, J, dual=True, device=device)
    W, WL = OperatorHandle(pack, "W"), OperatorHandle(pack, "WL")
    Pm, Pd = OperatorHandle(pack, "Pm"), OperatorHandle(pack, "Pd")
    mask, mask_lg = MaskHandle(pack, False), MaskHandle(pack, True)
    if not sparse:
        W, WL, Pm, Pd = (h.to_dense().cpu() for h in (W, WL, Pm, Pd))
        mask, mask_lg = mask.to_dense().cpu(), mask_lg.to_dense().cpu()
    return X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch
