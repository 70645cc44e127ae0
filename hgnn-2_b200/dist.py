"""Data-parallel plumbing: one process per GPU, graphs sharded across ranks, ONE flat fp32 gradient
buffer all-reduced per step over NCCL/NVLink, and a fused Adamax over the flat parameter buffer.

The reference is single-process (no distributed code at all); batches of graphs
are independent units (block-diagonal CSR has no cross-graph edges), so the path shards with no
data-path collective - the only exchange is the gradient sum.  The payload is a
few thousand floats (LGNN L=20, h=2: ~3 k parameters), i.e. latency-bound: one collective per step,
never per layer.  Batch-norm statistics stay per-rank (= the reference run on the local shard).
"""
import weakref

import torch
import torch.distributed as dist

from ._lib import call, fptr, stream

_FLAT_OF = weakref.WeakKeyDictionary()      # model -> weakref(FlatParams)


def flat_params_of(model):
    """The live ``FlatParams`` that re-homed this model's parameters, or None."""
    ref = _FLAT_OF.get(model)
    return ref() if ref is not None else None


def shard_range(n_items, rank, world):
    """Contiguous split of ``n_items`` units: rank r gets [lo, hi)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatParams(object):
    """Re-homes every parameter of ``model`` as a view into one flat buffer, and every ``.grad`` as a
    gathered into one flat gradient buffer (so a step needs one concat, one all-reduce, one optimizer
    launch).  Parameter names/shapes are untouched, so state_dicts still match the reference."""

    def __init__(self, model, fused_grad=True):
        """``fused_grad``: the model-level engine (engine.run_model) differentiates w.r.t. ONE leaf
        tensor aliasing the flat buffer (``flat_leaf``), so a step produces a single flat gradient
        instead of ~230 per-parameter views + AccumulateGrad nodes; per-parameter ``.grad`` views are
        then made on demand by ``scatter_grads()``.  False: gradients arrive per parameter."""
        self.params = [p for p in model.parameters()]
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.n = sum(sizes)
        self.flat = torch.empty(self.n, device=dev)
        self.grad = torch.zeros(self.n, device=dev)
        off = 0
        for p, k in zip(self.params, sizes):
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            off += k
        self.fused_grad = bool(fused_grad)
        self.fused_used = False
        self.flat_leaf = self.flat.detach().requires_grad_()     # same storage, its own autograd leaf
        self._last_off = off - sizes[-1]
        _FLAT_OF[model] = weakref.ref(self)

    def matches(self, params):
        """True while ``params`` are still the views this object created (cheap check on both ends)."""
        return (len(params) == len(self.params) and params[0] is self.params[0] and params[-1] is self.params[-1]
                and params[0].data_ptr() == self.flat.data_ptr()
                and params[-1].data_ptr() == self.flat.data_ptr() + 4 * self._last_off)

    def zero_grad(self):
        """Drop the gradients: autograd then *moves* each fresh gradient into place instead of
        launching one accumulate-add kernel per parameter (~230 per LGNN step)."""
        self.flat_leaf.grad = None
        if self.fused_used:
            self.fused_used = False
            if not self._scattered:
                return
        self._scattered = False
        for p in self.params:
            p.grad = None

    _scattered = False

    def scatter_grads(self):
        """Per-parameter ``.grad`` views of the flat gradient (fused_grad mode leaves them unset)."""
        g = self.flat_leaf.grad
        if g is None:
            return
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = g[off:off + k].view(p.shape)
            off += k
        self._scattered = True

    def _adopt_flat_grad(self):
        """The model-level engine returns every .grad as a view of ONE flat buffer laid out in
        parameter order: use it as is (no copy).  Returns False for any other layout."""
        if self.flat_leaf.grad is not None:
            self.grad = self.flat_leaf.grad
            return True
        g0 = self.params[0].grad
        if g0 is None or g0.dtype != torch.float32 or g0.storage_offset() != 0:
            return False
        store = g0.untyped_storage()
        if store.nbytes() < 4 * self.n:
            return False
        root = store.data_ptr()
        off = 0
        for p in self.params:
            g = p.grad
            if (g is None or g.untyped_storage().data_ptr() != root or g.storage_offset() != off
                    or not g.is_contiguous()):
                return False
            off += p.numel()
        base = torch.as_strided(g0, (self.n,), (1,), 0)
        self.grad = base
        return True

    def gather_grad(self):
        """Concatenate the per-parameter gradients into the flat buffer (one or two launches)."""
        if self._adopt_flat_grad():
            return
        parts = [p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel(), device=self.flat.device)
                 for p in self.params]
        torch.cat(parts, out=self.grad)

    def broadcast(self, src=0):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.flat, src)

    def all_reduce_grad(self):
        """Sum of the shard gradients; the 1/world factor is folded into the optimizer launch.  With a
        ``FusedAdamax(..., peer_allreduce=True)`` attached the sum happens INSIDE the optimizer launch (peer
        memory over NVLink, csrc/p2p.cu) and this only gathers the gradient."""
        self.gather_grad()
        if self.fused_allreduce:
            return
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)

    fused_allreduce = False


class FusedAdamax(object):
    """torch.optim.Adamax(lr) semantics (scripts/main_gnn.py:160-167) as one launch over the flat
    buffer (csrc/optim.cu).  The step counter lives on the device so a captured CUDA graph replays
    correctly.

    ``peer_allreduce`` (default: on when torch.distributed runs with more than one rank and the gradient fits the
    peer buffer; env HGNN_B200_NO_P2P=1 turns it off): the data-parallel gradient sum is fused into the optimizer
    launch - every rank publishes its flat gradient in a CUDA-IPC buffer, the kernel reads the peers' gradients
    over NVLink, sums them in rank order and updates (csrc/p2p.cu).  ``fp.all_reduce_grad()`` then does not call
    NCCL.  All ranks must construct the optimizer collectively (handle exchange) and step the same number of times."""

    def __init__(self, flat_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, peer_allreduce=None):
        import os
        self.fp, self.lr, self.betas, self.eps = flat_params, lr, betas, eps
        dev = flat_params.flat.device
        self.exp_avg = torch.zeros_like(flat_params.flat)
        self.exp_inf = torch.zeros_like(flat_params.flat)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.peers = None
        world = dist.get_world_size() if dist.is_initialized() else 1
        if peer_allreduce is None:
            peer_allreduce = world > 1 and os.environ.get("HGNN_B200_NO_P2P", "0") != "1"
        if peer_allreduce and world > 1:
            self._open_peers(dev, world)

    def _open_peers(self, dev, world):
        import ctypes
        from . import _lib
        n = self.fp.n
        ok = n <= int(_lib.lib.hgnn_p2p_max_floats()) and world <= 16
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if not all(flags):
            return                      # every rank falls back to NCCL together
        torch.cuda.set_device(dev)
        cap = int(n)
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        call("hgnn_p2p_alloc", cap, ctypes.byref(ptr), handle)
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw)
        rank = dist.get_rank()
        bufs = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                bufs[r] = ptr.value
            else:
                q = ctypes.c_void_p()
                call("hgnn_p2p_open", ctypes.create_string_buffer(handles[r], 64), ctypes.byref(q))
                bufs[r] = q.value
        dist.barrier()                  # every buffer exists and is zeroed before anybody's first step
        self.peers = (bufs, rank, world, cap, ptr.value)
        self.fault = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fp.fused_allreduce = True

    def step(self, grad_scale=1.0, collective=True):
        """One update.  ``collective=False`` (profiling only): the plain local update even when the peer path is on."""
        fp = self.fp
        if self.peers is not None and collective:
            bufs, rank, world, cap, _ = self.peers
            call("hgnn_p2p_allreduce_adamax", fptr(fp.flat), fptr(fp.grad), fptr(self.exp_avg), fptr(self.exp_inf),
                 fp.n, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                 float(grad_scale), self.step_count.data_ptr(), bufs, rank, world, cap, self.fault.data_ptr(), stream())
            return
        call("hgnn_adamax_step", fptr(fp.flat), fptr(fp.grad), fptr(self.exp_avg), fptr(self.exp_inf),
             fp.n, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
             float(grad_scale), self.step_count.data_ptr(), stream())
