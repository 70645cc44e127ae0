"""Model-level training engine for ``GNN_simple`` / ``GNN_lg`` (csrc/engine.cu).

``models.gnns.model_mnb`` keeps the reference's classes and signatures; their ``forward`` hands the
whole layer stack to ``run_model`` below, which executes it as ONE ``torch.autograd.Function``:

* forward: one fused launch per layer side (``hgnn_lg_side_fwd``).  Activations stay RAW
  (pre-batch-norm) in HBM; each consumer normalises its inputs on load from the producer's fp64
  (sum z, sum z^2) accumulators, so there is no BN-apply pass and no finalisation tail;
* backward: one launch per layer side (``hgnn_lg_side_bwd``) - BN + ReLU backward on the fly while
  gathering through the transposed operators, both input gradients, dW / dbias and the BN sums of
  the produced gradients.  Input gradients accumulate in place (no torch ``add`` kernels);
* step end: ``hgnn_bins_reduce`` converts every binned fp64 accumulator of the step into one flat
  fp32 gradient buffer in ``model.parameters()`` order (the per-parameter ``.grad`` tensors are views
  of it, which ``dist.FlatParams`` recognises), ``hgnn_bn_running_update`` applies the
  running-statistics rule for all BN instances at once.

Semantics = the reference's layer stack (models/gnns/model_mnb.py:58-66,124-129 over
models/layers/layers_mnb.py); the layer-level modules keep their own (per-module) kernels for the
reference's layer API.
"""
import ctypes
import os
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import (BatchT, BnRefT, ProgramT, ProgSideT, ProgTensorT, SideBwdT, SideT, call, call_program, fptr, iptr,
                   make_ops, stream)

# Training steps run through the native program executor (csrc/program.cu: the side loop in C++, two
# foreign calls per step).  False (or env HGNN_B200_PY_ENGINE=1): the per-side Python loop below, which
# issues the same launches one ctypes call at a time - kept for eval mode, SPLIT_DW and bench.py's
# per-entry-point profiling.
USE_PROGRAM = os.environ.get("HGNN_B200_PY_ENGINE", "0") != "1"


# Optional: split the weight gradients of width-4 sides off the backward gather chain (x1 rows saved
# by the forward, streaming hgnn_lg_side_dw on a parallel stream).  Measured on the C2 workload it
# shortens the gather kernels (21.7 -> 17.7 us, 16.9 -> 13.3 us) but the extra launches share the SMs
# with the chain and the step does not get faster (1.38 vs 1.30 ms), so it is off by default; kept,
# tested, for the multi-side variant planned in profiles/README.md.
SPLIT_DW = False
_side_streams = {}


def _side_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


_MEGA_SCRATCH = {}


_MEGA_ON = os.environ.get("HGNN_B200_MEGA", "0") == "1"


class _NoScratch(object):
    @staticmethod
    def data_ptr():
        return None


def _mega_scratch(device):
    """256 zeroed bytes per (device, stream): the grid-barrier state of the persistent kernels (csrc/mega.cu).
    Only allocated when those kernels are switched on (HGNN_B200_MEGA=1)."""
    if not _MEGA_ON:
        return _NoScratch
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, stream())
    buf = _MEGA_SCRATCH.get(key)
    if buf is None:
        buf = _MEGA_SCRATCH[key] = torch.zeros(64, dtype=torch.int32, device=torch.device("cuda", idx))
    return buf


def _bins(width):
    return int(_lib.lib.hgnn_bins_for(int(width)))


class _Side(object):
    """One fused side update: out = BN(cat(cv_a(x1), relu(cv_b(x1)))), x1 = [ops(self) | Pm/Pd(cross)]."""
    __slots__ = ("name", "kind", "src_self", "src_cross", "out", "conv_a", "conv_b", "bn", "relu_from",
                 "Fs", "Fc", "Fout", "Cin", "dW_off", "db_off")


class _Plan(object):
    """Static execution plan of a model: sides in forward order, tensor table, arena layout and the
    accumulator -> flat-gradient table."""

    def __init__(self, model):
        lg = bool(getattr(model, "dual", False))
        order = getattr(model, "order", 0) if lg else 0
        K = model.J + 2
        h = model.n_features
        self.lg, self.order, self.K = lg, order, K
        layers = [model.layer0] + [model._modules["layer{}".format(i + 1)] for i in range(model.n_layers - 2)]
        self.sides = []
        self.tensors = {"X": dict(F=model.featuremap_in[0], rows="n", bn=None)}
        if lg:
            self.tensors["XL"] = dict(F=1, rows="m", bn=None)
        cur_n, cur_e = "X", "XL"

        def add(name, kind, src_self, src_cross, out, conv_a, conv_b, bn, relu_from, rows):
            s = _Side()
            s.name, s.kind, s.src_self, s.src_cross, s.out = name, kind, src_self, src_cross, out
            s.conv_a, s.conv_b, s.bn, s.relu_from = conv_a, conv_b, bn, relu_from
            s.Fs = self.tensors[src_self]["F"]
            s.Fc = self.tensors[src_cross]["F"] if src_cross else 0
            s.Fout = conv_a.weight.shape[0] + (conv_b.weight.shape[0] if conv_b is not None else 0)
            s.Cin = K * s.Fs + 2 * s.Fc
            assert conv_a.weight.shape[1] == s.Cin, (name, conv_a.weight.shape, s.Cin)
            if out is not None:
                self.tensors[out] = dict(F=s.Fout, rows=rows, bn=bn)
            self.sides.append(s)

        for li, layer in enumerate(layers):
            n_out, e_out = "N%d" % li, "E%d" % li
            if not lg:
                add("L%d.node" % li, "node", cur_n, None, n_out, layer.cv2, layer.cv1, layer.bn1, 0, "n")
                cur_n = n_out
                continue
            node = lambda cross: add("L%d.node" % li, "node", cur_n, cross, n_out, layer.cv2, layer.cv1,  # noqa: E731
                                     layer.bn1, h, "n")
            edge = lambda cross: add("L%d.edge" % li, "edge", cur_e, cross, e_out, layer.cv4, layer.cv3,  # noqa: E731
                                     layer.bn2, h, "m")
            if order == 1:
                node(cur_e)
                edge(n_out)
            elif order == 2:
                edge(cur_n)
                node(e_out)
            else:
                node(cur_e)
                edge(cur_n)
            cur_n, cur_e = n_out, e_out
        fc = model.layerlast.fc
        add("readout", "node", cur_n, cur_e if lg else None, None, fc, None, None, fc.weight.shape[0], "n")

        # ---- arena layout (fp64): per BN'd tensor acc_f | acc_b ; per side dW | db
        off = 0
        for name, t in self.tensors.items():
            if t["bn"] is not None:
                w = 2 * t["F"]
                t["acc_f"], t["acc_b"] = off, off + _bins(w) * w
                off += 2 * _bins(w) * w
        for s in self.sides:
            s.dW_off = off
            off += _bins(s.Fout * s.Cin) * s.Fout * s.Cin
            s.db_off = off
            off += _bins(s.Fout) * s.Fout
        self.arena_size = off

        # ---- accumulator -> flat gradient table, in model.parameters() order
        loc = {}
        for s in self.sides:
            Ha = s.conv_a.weight.shape[0]
            wn, bnn = s.Fout * s.Cin, s.Fout
            for conv, o0 in ((s.conv_a, 0), (s.conv_b, Ha)):
                if conv is None:
                    continue
                n = conv.weight.numel()
                loc[id(conv.weight)] = (s.dW_off + o0 * s.Cin + np.arange(n), _bins(wn), wn, 1)
                loc[id(conv.bias)] = (s.db_off + o0 + np.arange(conv.bias.numel()), _bins(bnn), bnn, 1)
        for name, t in self.tensors.items():
            if t["bn"] is not None:
                F = t["F"]
                loc[id(t["bn"].weight)] = (np.array([t["acc_b"] + F]), _bins(2 * F), 2 * F, F)
                loc[id(t["bn"].bias)] = (np.array([t["acc_b"]]), _bins(2 * F), 2 * F, F)
        offs, nbs, strides, cnts = [], [], [], []
        self.params = list(model.parameters())
        self.param_slices = []
        pos = 0
        for p in self.params:
            n = p.numel()
            self.param_slices.append((pos, n, tuple(p.shape)))
            pos += n
            if id(p) in loc:
                o, nb, st, cnt = loc[id(p)]
                offs.append(o.astype(np.int64))
                nbs.append(np.full(n, nb, np.int32))
                strides.append(np.full(n, st, np.int32))
                cnts.append(np.full(n, cnt, np.int32))
            else:     # parameter off the hot path (GRUUpdate): zero gradient
                offs.append(np.zeros(n, np.int64))
                nbs.append(np.zeros(n, np.int32))
                strides.append(np.zeros(n, np.int32))
                cnts.append(np.zeros(n, np.int32))
        self.n_flat = pos
        self.table_host = (np.concatenate(offs), np.concatenate(nbs), np.concatenate(strides), np.concatenate(cnts))
        self.bn_list = [t for t in self.tensors.values() if t["bn"] is not None]
        self._device_tables = {}
        self._running = None
        self._programs = {}
        self._addr_cache = None
        self.fits = None          # filled by supported(): every side inside the kernels' shared-memory budgets
        self.readout_width = self.sides[-1].Fout

        # ---- the same plan as C structs for csrc/program.cu (parameters by index)
        pidx = {id(p): i for i, p in enumerate(self.params)}
        tidx = {name: i for i, name in enumerate(self.tensors)}
        self.c_tensors = (ProgTensorT * len(self.tensors))()
        for i, t in enumerate(self.tensors.values()):
            ct = self.c_tensors[i]
            ct.F, ct.rows = t["F"], 1 if t["rows"] == "m" else 0
            if t["bn"] is None:
                ct.bn_weight = ct.bn_bias = -1
                ct.acc_f = ct.acc_b = 0
            else:
                ct.bn_weight, ct.bn_bias = pidx[id(t["bn"].weight)], pidx[id(t["bn"].bias)]
                ct.acc_f, ct.acc_b = t["acc_f"], t["acc_b"]
        self.c_sides = (ProgSideT * len(self.sides))()
        for i, sd in enumerate(self.sides):
            cs = self.c_sides[i]
            cs.kind = 0 if sd.kind == "node" else 1
            cs.src_self = tidx[sd.src_self]
            cs.src_cross = tidx[sd.src_cross] if sd.src_cross else -1
            cs.out = tidx[sd.out] if sd.out is not None else -1
            cs.Wa, cs.ba, cs.Ha = pidx[id(sd.conv_a.weight)], pidx[id(sd.conv_a.bias)], sd.conv_a.weight.shape[0]
            if sd.conv_b is not None:
                cs.Wb, cs.bb, cs.Hb = pidx[id(sd.conv_b.weight)], pidx[id(sd.conv_b.bias)], sd.conv_b.weight.shape[0]
            else:
                cs.Wb, cs.bb, cs.Hb = -1, -1, 0
            cs.relu_from, cs.dW_off, cs.db_off = sd.relu_from, sd.dW_off, sd.db_off

    def program(self, device):
        """``ProgramT`` for csrc/program.cu with this device's tables (rebuilt when the running
        statistics were re-homed)."""
        flat, tabs = self.running_flat(device)
        key = str(device)
        hit = self._programs.get(key)
        if hit is not None and hit[1] is tabs:
            return hit[0]
        offs, nbs, strides, cnts = self.tables(device)
        pr = ProgramT()
        pr.n_tensors, pr.tensors = len(self.tensors), self.c_tensors
        pr.n_sides, pr.sides = len(self.sides), self.c_sides
        pr.dual, pr.arena_doubles = 1 if self.lg else 0, self.arena_size
        pr.n_flat = self.n_flat
        pr.red_off, pr.red_nb, pr.red_stride, pr.red_cnt = (offs.data_ptr(), nbs.data_ptr(), strides.data_ptr(),
                                                            cnts.data_ptr())
        keep = None
        if tabs is not None:
            acc_off, Fs, run_off = tabs
            kinds = torch.tensor([1 if t["rows"] == "m" else 0 for t in self.bn_list], dtype=torch.int32,
                                 device=device)
            pr.n_bn = len(self.bn_list)
            pr.bn_acc_off, pr.bn_F, pr.bn_rows_kind, pr.bn_run_off = (acc_off.data_ptr(), Fs.data_ptr(),
                                                                      kinds.data_ptr(), run_off.data_ptr())
            pr.momentum = float(self.bn_list[0]["bn"].momentum)
            keep = kinds
        else:
            pr.n_bn = 0
        self._programs[key] = (pr, tabs, keep)
        return pr

    def sync_bn_scale(self, device, world):
        """Flat vector: 1 / world at the positions of batch-norm weights and biases, 1 elsewhere (sync_bn)."""
        key = (str(device), world)
        hit = getattr(self, "_sync_scale", None)
        if hit is None or hit[0] != key:
            v = torch.ones(self.n_flat)
            bn_ids = {id(t["bn"].weight) for t in self.bn_list} | {id(t["bn"].bias) for t in self.bn_list}
            for p_, (pos, n, _shape) in zip(self.params, self.param_slices):
                if id(p_) in bn_ids:
                    v[pos:pos + n] = 1.0 / world
            hit = self._sync_scale = (key, v.to(device))
        return hit[1]

    def param_addresses(self):
        """Device addresses of all parameters (numpy int64, model.parameters() order).  Layout and
        dtype are validated on the first and last parameter per call, on all of them when the plan
        is built (``_param_ptrs``).  The table is rebuilt only when one of those two addresses moved
        (walking ~230 parameters costs 0.1 ms - as much as issuing a third of a pass)."""
        p0, p1 = fptr(self.params[0]), fptr(self.params[-1])
        hit = self._addr_cache
        if hit is not None and hit[0] == p0 and hit[1] == p1:
            return hit[2]
        addr = np.fromiter((p.data_ptr() for p in self.params), dtype=np.int64, count=len(self.params))
        self._addr_cache = (p0, p1, addr)
        return addr

    # ---- per-device tables ---------------------------------------------------------------------
    def tables(self, device):
        key = str(device)
        if key not in self._device_tables:
            o, nb, st, cnt = self.table_host
            self._device_tables[key] = (torch.from_numpy(o).to(device), torch.from_numpy(nb).to(device),
                                        torch.from_numpy(st).to(device), torch.from_numpy(cnt).to(device))
        return self._device_tables[key]

    def running_flat(self, device):
        """All running_mean / running_std buffers as views of one flat buffer (re-homed lazily, and
        again whenever the module was moved and its buffers were replaced)."""
        first = self.bn_list[0]["bn"] if self.bn_list else None
        if first is None:
            return None, None
        r = self._running
        if r is not None and r[0].device == device and first.running_mean.data_ptr() == r[0].data_ptr():
            return r
        total = sum(2 * t["F"] for t in self.bn_list)
        flat = torch.zeros(total, device=device)
        run_off, acc_off, Fs = [], [], []
        pos = 0
        for t in self.bn_list:
            bn, F = t["bn"], t["F"]
            flat[pos:pos + F].copy_(bn.running_mean.to(device))
            flat[pos + F:pos + 2 * F].copy_(bn.running_std.to(device))
            bn.running_mean = flat[pos:pos + F]
            bn.running_std = flat[pos + F:pos + 2 * F]
            run_off.append(pos)
            acc_off.append(t["acc_f"])
            Fs.append(F)
            pos += 2 * F
        tabs = (torch.tensor(acc_off, dtype=torch.int64, device=device),
                torch.tensor(Fs, dtype=torch.int32, device=device),
                torch.tensor(run_off, dtype=torch.int64, device=device))
        self._running = (flat, tabs)
        return self._running


# model -> _Plan.  Kept OUTSIDE the module's __dict__: a plan holds ctypes structs with raw device
# pointers, which neither pickle (whole-module torch.save, functions/logs.py:99-111 of the reference)
# nor copy.deepcopy can serialise; a reloaded / copied model simply gets a fresh plan on first use.
_PLANS = weakref.WeakKeyDictionary()


def get_plan(model):
    """The cached plan; rebuilt when the parameter objects were replaced (checked on two of them -
    walking all ~230 parameters costs more than a kernel launch)."""
    plan = _PLANS.get(model)
    if (plan is None or plan.params[0] is not model.layer0.cv1.weight
            or plan.params[-1] is not model.layerlast.fc.bias):
        model.__dict__.pop("_engine_plan", None)      # attribute written by older versions of this file
        plan = _Plan(model)
        _PLANS[model] = plan
    return plan


def _rows_table(pack, plan, device):
    cache = pack.__dict__.setdefault("_engine_rows", {})
    key = id(plan)
    if key not in cache:
        rows = [pack.Rn if t["rows"] == "n" else pack.Rm for t in plan.bn_list]
        cache[key] = torch.tensor(rows, dtype=torch.int32, device=device)
    return cache[key]


def _pack_cache(pack):
    """ctypes operator arrays and raw pointers of a pack, built once (host overhead matters: an eager
    step is ~90 C-ABI calls)."""
    c = pack.__dict__.get("_engine_cache")
    if c is None:
        c = {"node": make_ops(pack.node_ops()), "nodeT": make_ops(pack.node_ops_T())}
        if pack.dual:
            c["edge"] = make_ops(pack.edge_ops())
            c["edgeT"] = make_ops(pack.edge_ops_T(split=True))
            for k, p in (("p", pack.p), ("pt", pack.pt)):
                c[k] = (iptr(p.rowptr), iptr(p.col), fptr(p.val), fptr(p.val2))
            c["p_nnz"] = pack.p.nnz
        c["node_off"], c["pad_n"] = iptr(pack.node_off), fptr(pack.pad_n)
        b = BatchT()
        b.bs, b.Rn, b.Rm, b.n_ops = pack.bs, pack.Rn, (pack.Rm if pack.dual else 0), pack.K
        b.node_ops, b.node_ops_T = c["node"][0], c["nodeT"][0]
        if pack.dual:
            b.edge_ops, b.edge_ops_T = c["edge"][0], c["edgeT"][0]
            b.p_rowptr, b.p_col, b.p_pm, b.p_pd = c["p"]
            b.pt_rowptr, b.pt_col, b.pt_pm, b.pt_pd = c["pt"]
            b.p_nnz = c["p_nnz"]
        b.node_off, b.pad_n = c["node_off"], c["pad_n"]
        # persistent-kernel path (csrc/mega.cu): collapsed line graph + grid-barrier scratch
        b.collapse_ok = 0
        b.mega_scratch = _mega_scratch(pack.device).data_ptr()
        if pack.dual and not pack.generic and getattr(pack, "btc", None) is not None:
            c["btc"] = (iptr(pack.btc.rowptr), iptr(pack.btc.col), fptr(pack.btc.val), iptr(pack.erow), fptr(pack.ew))
            c["edgeTc"] = make_ops([("ident",), ("diag", pack.dl), pack.btc.desc()])     # [I, D, btc]
            c["edgeFc"] = make_ops(pack.edge_ops())       # forward operators (same hints as the executor: csrc/program.cu)
            b.btc_rowptr, b.btc_col, b.btc_val, b.erow, b.ew = c["btc"]
            b.n_act = int(pack.erow.numel())
            b.btc_nnz = int(pack.btc.nnz)
        c["batch"] = b
        pack.__dict__["_engine_cache"] = c
    return c


def _uses_collapse(plan, pack, dev, xl_is_degree):
    """Whether the step executor runs this (model, batch) on the collapsed line graph (asked of the library, which
    decides: ``hgnn_program_uses_collapse``); the per-side Python loop below then does the same."""
    if not (xl_is_degree and plan.lg and "btc" in _pack_cache(pack)):
        return False
    batch = _pack_cache(pack)["batch"]
    batch.collapse_ok = 1
    return _lib.lib.hgnn_program_uses_collapse(ctypes.byref(plan.program(dev)), ctypes.byref(batch)) == 1


def _side_struct(pack, side, Xs, Xc, collapse=False):
    node = side.kind == "node"
    pc = _pack_cache(pack)
    ops, n = pc["node" if node else "edge"]
    s = SideT()
    s.R, s.ops, s.n_ops = (pack.Rn if node else pack.Rm), ops, n
    if collapse and not node:      # only the active line-graph rows, the representative weighted by its multiplicity
        s.rowmap, s.roww, s.R = pc["btc"][3], pc["btc"][4], int(pack.erow.numel())
        ops, n = pc["edgeFc"]
        s.ops = ops
    s.Xs, s.Fs = Xs.data_ptr(), Xs.shape[1]
    if Xc is not None:
        s.p_rowptr, s.p_col, s.p_pm, s.p_pd = pc["p" if node else "pt"]
        s.Xc, s.Fc = Xc.data_ptr(), Xc.shape[1]
        s.p_nnz = pc["p_nnz"]
    else:
        s.p_rowptr = s.p_col = s.p_pm = s.p_pd = s.Xc = None
        s.Fc = 0
    return s, ops


def _param_ptrs(plan):
    """{id(parameter): device pointer}, validated once per call (CUDA, fp32, contiguous)."""
    out = {}
    for p in plan.params:
        out[id(p)] = fptr(p)
    return out


def _bn_ref(plan, tname, arena, rows, affine=None, pp=None):
    t = plan.tensors[tname]
    r = BnRefT()
    if t["bn"] is None:
        r.acc = r.affine = r.weight = r.bias = None
        r.n_rows = 0
        return r
    r.weight, r.bias, r.n_rows = pp[id(t["bn"].weight)], pp[id(t["bn"].bias)], rows
    if affine is not None:
        r.acc, r.affine = None, affine[tname].data_ptr()
    else:
        r.acc, r.affine = arena.data_ptr() + 8 * t["acc_f"], None
    return r


def _rows_of(pack, plan, tname, glob=None):
    """Rows behind a tensor's batch-norm statistics: the local rows, or with ``sync_bn`` the rows of all ranks."""
    node = plan.tensors[tname]["rows"] == "n"
    if glob is not None:
        return glob[0] if node else glob[1]
    return pack.Rn if node else pack.Rm


def _sync_world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _global_rows(pack, dev):
    """(sum of Rn, sum of Rm) over all ranks - the divisors of synchronised batch-norm (one tiny all-reduce per batch)."""
    import torch.distributed as dist
    t = torch.tensor([pack.Rn, pack.Rm], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    n, m = t.tolist()
    return int(n), int(m)


def _allreduce_bins(arena, off_doubles, width):
    """Sum one binned accumulator block of the step arena over all ranks, in place (sync_bn)."""
    import torch.distributed as dist
    n = _bins(width) * width
    dist.all_reduce(arena[off_doubles:off_doubles + n])


def _forward(plan, pack, Xp, XLp, training, arena, save_x1=False, collapse=False, glob=None):
    """Runs every side; returns (dict of raw tensors, model output).  ``glob`` = (global Rn, global Rm): synchronised
    batch-norm - the (sum z, sum z^2) bins of every side are summed over the ranks before its consumers run."""
    dev = Xp.device
    vals = {"X": Xp, "XL": XLp}
    pp = _param_ptrs(plan)
    pc = _pack_cache(pack)
    cur = stream()
    arena_ptr = arena.data_ptr()
    affine = None
    if not training:      # eval: normalise with the running statistics (batch_normalization.py:39-41)
        affine = {}
        for name, t in plan.tensors.items():
            if t["bn"] is not None:
                bn, F = t["bn"], t["F"]
                st = torch.empty(4 * F, device=dev)
                call("hgnn_bn_stats_eval", F, fptr(bn.weight), fptr(bn.bias), fptr(bn.running_mean),
                     fptr(bn.running_std), fptr(st), stream())
                affine[name] = st[2 * F:]
    out = None
    for s in plan.sides:
        Xs = vals[s.src_self]
        Xc = vals[s.src_cross] if s.src_cross else None
        st, keep = _side_struct(pack, s, Xs, Xc, collapse)
        bs_ = _bn_ref(plan, s.src_self, arena, _rows_of(pack, plan, s.src_self, glob), affine, pp)
        bc_ = _bn_ref(plan, s.src_cross, arena, _rows_of(pack, plan, s.src_cross, glob), affine, pp) if s.src_cross else None
        Ha = s.conv_a.weight.shape[0]
        Hb = s.conv_b.weight.shape[0] if s.conv_b is not None else 0
        Z = torch.empty(pack.Rn if s.kind == "node" else pack.Rm, s.Fout, device=dev)   # all rows (st.R may be the active rows only)
        _lib.tag = s.name
        acc_out = None
        if s.out is not None and training:
            acc_out = arena_ptr + 8 * plan.tensors[s.out]["acc_f"]
        # width-4 fast path in training: save the concatenated x1 rows so that the weight gradients
        # become a streaming pass on a parallel branch (hgnn_lg_side_dw) instead of 48+32 register
        # accumulators inside the latency-bound backward gather
        X1 = None
        if save_x1 and s.out is not None and _lib.lib.hgnn_lg_row4_eligible(keep, st.n_ops, s.Fs, s.Fc, s.Fout):
            X1 = torch.empty(pack.Rn if s.kind == "node" else pack.Rm, s.Cin, device=dev)
            vals["x1:" + s.name] = X1
        call("hgnn_lg_side_fwd", ctypes.byref(st), ctypes.byref(bs_), ctypes.byref(bc_) if bc_ is not None else None,
             pp[id(s.conv_a.weight)], pp[id(s.conv_a.bias)], Ha,
             pp[id(s.conv_b.weight)] if Hb else None, pp[id(s.conv_b.bias)] if Hb else None, Hb,
             s.relu_from, Z.data_ptr(), acc_out, X1.data_ptr() if X1 is not None else None, cur)
        if s.out is not None:
            vals[s.out] = Z
            if glob is not None and training:
                _allreduce_bins(arena, plan.tensors[s.out]["acc_f"], 2 * s.Fout)
        else:     # readout: sum over all Nmax slots, padded slots add fc.bias (layers_mnb.py:92,:386)
            out = torch.empty(pack.bs, s.Fout, device=dev)
            call("hgnn_segment_sum", Z.data_ptr(), pack.bs, s.Fout, pc["node_off"], pc["pad_n"],
                 pp[id(s.conv_a.bias)], out.data_ptr(), cur)
    return vals, out


class _ModelFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, model, pack, Xp, XLp, flat_mode, xl_is_degree, *params):
        plan = get_plan(model)
        dev = Xp.device
        ctx.flat_mode = flat_mode
        ctx.xl_is_degree = 1 if xl_is_degree else 0
        # synchronised batch-norm (model.sync_bn = True under torch.distributed): statistics over the batches of ALL
        # ranks, as one process on the global batch would compute them (batch_normalization.py:80-93).
        # It needs a collective between consecutive sides, so the per-side Python loop runs instead of the executor.
        ctx.glob = _global_rows(pack, dev) if (getattr(model, "sync_bn", False) and _sync_world() > 1) else None
        ctx.program = USE_PROGRAM and not SPLIT_DW and ctx.glob is None
        if ctx.program:
            prog = plan.program(dev)
            batch = _pack_cache(pack)["batch"]
            n_work = int(_lib.lib.hgnn_program_work_floats(ctypes.byref(prog), batch.Rn, batch.Rm))
            work = torch.empty(max(n_work, 1), device=dev)
            arena = torch.empty(max(plan.arena_size, 1), dtype=torch.float64, device=dev)
            out = torch.empty(pack.bs, plan.readout_width, device=dev)
            addr = plan.param_addresses()
            run = plan.running_flat(dev)[0]
            _lib.tag = "program"
            batch.collapse_ok = ctx.xl_is_degree
            batch.mega_scratch = _mega_scratch(dev).data_ptr()
            call_program("hgnn_program_fwd", ctypes.byref(prog), ctypes.byref(batch), fptr(Xp),
                         fptr(XLp) if XLp is not None else None, addr.ctypes.data, work.data_ptr(), arena.data_ptr(),
                         run.data_ptr() if run is not None else None, out.data_ptr(), stream())
            ctx.plan, ctx.pack, ctx.arena, ctx.work, ctx.inputs = plan, pack, arena, work, (Xp, XLp)
            ctx.need_x = Xp.requires_grad
            return out
        arena = torch.zeros(max(plan.arena_size, 1), dtype=torch.float64, device=dev)
        ctx.collapse = _uses_collapse(plan, pack, dev, xl_is_degree) and not SPLIT_DW
        vals, out = _forward(plan, pack, Xp, XLp, True, arena, save_x1=SPLIT_DW and any(ctx.needs_input_grad),
                             collapse=ctx.collapse, glob=ctx.glob)
        run = plan.running_flat(dev)
        if run[0] is not None:
            flat, (acc_off, Fs, run_off) = run
            mom = float(plan.bn_list[0]["bn"].momentum)
            if ctx.glob is not None:
                rows_t = torch.tensor([ctx.glob[0] if t["rows"] == "n" else ctx.glob[1] for t in plan.bn_list],
                                      dtype=torch.int32, device=dev)
            else:
                rows_t = _rows_table(pack, plan, dev)
            call("hgnn_bn_running_update", arena.data_ptr(), acc_off.data_ptr(), iptr(Fs),
                 iptr(rows_t), run_off.data_ptr(), len(plan.bn_list), mom, fptr(flat), stream())
        ctx.plan, ctx.pack, ctx.vals, ctx.arena = plan, pack, vals, arena
        ctx.need_x = Xp.requires_grad
        return out

    @staticmethod
    def _grads_out(ctx, plan, gX, gflat):
        if ctx.flat_mode:
            return (None, None, gX, None, None, None, gflat)
        return (None, None, gX, None, None, None) + tuple(gflat[o:o + n].view(shape) for o, n, shape in plan.param_slices)

    @staticmethod
    def backward(ctx, g_out):
        plan, pack, arena = ctx.plan, ctx.pack, ctx.arena
        dev = g_out.device
        g_out = g_out.contiguous().float()
        if ctx.program:
            prog = plan.program(dev)
            batch = _pack_cache(pack)["batch"]
            Xp, XLp = ctx.inputs
            gwork = torch.empty_like(ctx.work)
            gflat = torch.empty(plan.n_flat, device=dev)
            n_scr = int(_lib.lib.hgnn_program_rng_scratch_bytes(ctypes.byref(prog), ctypes.byref(batch)))
            scratch = torch.empty(n_scr, dtype=torch.uint8, device=dev) if n_scr > 0 else None
            gX = torch.empty_like(Xp) if ctx.need_x else None
            addr = plan.param_addresses()
            batch.collapse_ok = ctx.xl_is_degree
            batch.mega_scratch = _mega_scratch(dev).data_ptr()
            call_program("hgnn_program_bwd", ctypes.byref(prog), ctypes.byref(batch), fptr(Xp),
                         fptr(XLp) if XLp is not None else None, addr.ctypes.data, ctx.work.data_ptr(),
                         gwork.data_ptr(), arena.data_ptr(), fptr(g_out), gX.data_ptr() if gX is not None else None,
                         gflat.data_ptr(), scratch.data_ptr() if scratch is not None else None, n_scr, stream())
            return _ModelFunction._grads_out(ctx, plan, gX, gflat)
        vals = ctx.vals
        glob = getattr(ctx, "glob", None)
        grads, started = {}, set()
        base = arena.data_ptr()
        main, side_stream, forked = torch.cuda.current_stream(), None, False
        pp = _param_ptrs(plan)
        pc = _pack_cache(pack)
        cur = stream()

        def grad_buf(name):
            if name not in grads:
                grads[name] = torch.empty_like(vals[name])
            return grads[name]

        # zeroed scratch for the dedicated range-sum CTAs of the width-4 edge-side backward (one region per side)
        rng_bytes, rng_scratch = 0, None
        collapse = getattr(ctx, "collapse", False)
        if pack.dual and not pack.generic and getattr(pack, "bts", None) is not None and not collapse:
            rng_bytes = int(_lib.lib.hgnn_lg_rng_scratch_bytes(int(pack.bts_ranges[3].numel())))
            if rng_bytes:
                rng_scratch = torch.zeros(rng_bytes * len(plan.sides), dtype=torch.uint8, device=dev)
        for si, s in reversed(list(enumerate(plan.sides))):
            node = s.kind == "node"
            d = SideBwdT()
            if rng_scratch is not None and not node:
                d.rng_scratch = rng_scratch.data_ptr() + si * rng_bytes
            Ha = s.conv_a.weight.shape[0]
            Hb = s.conv_b.weight.shape[0] if s.conv_b is not None else 0
            if s.out is None:       # readout: gPre = g_out broadcast over the rows of each graph
                G = torch.empty(pack.Rn, s.Fout, device=dev)
                call("hgnn_readout_bwd_prep", fptr(g_out), pack.bs, s.Fout, pc["node_off"], pc["pad_n"],
                     G.data_ptr(), base + 8 * s.db_off, cur)
                d.gY, d.Z, d.acc_f, d.acc_b, d.bn_weight = G.data_ptr(), None, None, None, None
                d.Rg = pack.Rn
                keep_g = G
            else:
                t = plan.tensors[s.out]
                if s.out not in grads:          # output never used downstream: zero gradient
                    grads[s.out] = torch.zeros_like(vals[s.out])
                d.gY, d.Z = grads[s.out].data_ptr(), vals[s.out].data_ptr()
                d.acc_f, d.acc_b = base + 8 * t["acc_f"], base + 8 * t["acc_b"]
                d.bn_weight = pp[id(t["bn"].weight)]
                d.Rg = _rows_of(pack, plan, s.out, glob)
                if glob is not None:      # every contribution to (sum g, sum g xhat) of this tensor is in: make it global
                    _allreduce_bins(arena, t["acc_b"], 2 * t["F"])
            d.Fg, d.relu_from = s.Fout, s.relu_from
            d.Wa, d.Ha = pp[id(s.conv_a.weight)], Ha
            d.Wb, d.Hb = (pp[id(s.conv_b.weight)] if Hb else None), Hb
            d.Cin = s.Cin
            d.dW_bins, d.db_bins = base + 8 * s.dW_off, base + 8 * s.db_off
            # self part
            opsT, n = pc["nodeT" if node else ("edgeTc" if collapse else "edgeT")]
            d.R_self, d.ops_T, d.n_ops = (pack.Rn if node else pack.Rm), opsT, n
            if collapse and not node:
                d.rowmap_self, d.roww_self, d.R_self = pc["btc"][3], pc["btc"][4], int(pack.erow.numel())
            d.Xs, d.Fs = vals[s.src_self].data_ptr(), s.Fs
            d.bn_self = _bn_ref(plan, s.src_self, arena, _rows_of(pack, plan, s.src_self, glob), None, pp)
            ts = plan.tensors[s.src_self]
            need_self = ts["bn"] is not None or (s.src_self == "X" and ctx.need_x)
            d.gXs = grad_buf(s.src_self).data_ptr() if need_self else None
            d.accumulate_self = 1 if s.src_self in started else 0
            d.acc_b_self = (base + 8 * ts["acc_b"]) if ts["bn"] is not None else None
            if need_self:
                started.add(s.src_self)
            # cross part
            if s.src_cross:
                tc = plan.tensors[s.src_cross]
                d.R_cross = _rows_of(pack, plan, s.src_cross)
                d.pt_rowptr, d.pt_col, d.pt_pm, d.pt_pd = pc["pt" if node else "p"]   # rows = the cross tensor's rows
                d.Xc, d.Fc = vals[s.src_cross].data_ptr(), s.Fc
                d.pt_nnz = pc["p_nnz"]
                d.bn_cross = _bn_ref(plan, s.src_cross, arena, _rows_of(pack, plan, s.src_cross, glob), None, pp)
                if collapse and node:       # the cross rows of a node side are line-graph rows
                    d.rowmap_cross, d.roww_cross, d.R_cross = pc["btc"][3], pc["btc"][4], int(pack.erow.numel())
                need_cross = tc["bn"] is not None or (s.src_cross == "X" and ctx.need_x)
                d.gXc = grad_buf(s.src_cross).data_ptr() if need_cross else None
                d.accumulate_cross = 1 if s.src_cross in started else 0
                d.acc_b_cross = (base + 8 * tc["acc_b"]) if tc["bn"] is not None else None
                if need_cross:
                    started.add(s.src_cross)
            else:
                d.R_cross = 0
            _lib.tag = s.name
            X1 = vals.get("x1:" + s.name)
            d.skip_dw = 0
            if X1 is not None:
                # gY of this side is complete here: fork the streaming dW pass onto the side stream
                d.skip_dw = 1
                if side_stream is None:
                    side_stream = _side_stream(dev)
                ev = torch.cuda.Event()
                ev.record(main)
                side_stream.wait_event(ev)
                with torch.cuda.stream(side_stream):
                    call("hgnn_lg_side_dw", d.gY, d.Z, d.Rg, s.relu_from, d.acc_f, d.acc_b, d.bn_weight,
                         fptr(X1), s.Cin, base + 8 * s.dW_off, base + 8 * s.db_off, stream())
                forked = True
            call("hgnn_lg_side_bwd", ctypes.byref(d), cur)
        if forked:
            main.wait_stream(side_stream)       # join before the accumulators are read
        offs, nbs, strides, cnts = plan.tables(dev)
        gflat = torch.empty(plan.n_flat, device=dev)
        call("hgnn_bins_reduce", arena.data_ptr(), offs.data_ptr(), iptr(nbs), iptr(strides), iptr(cnts),
             plan.n_flat, fptr(gflat), stream())
        if glob is not None:
            # The batch-norm weight / bias gradients were formed from GLOBAL sums: every rank already holds the sum
            # over ranks.  Pre-divide them by the world size so that the flat-gradient sum over ranks (followed by the
            # 1 / world of the optimizer) treats them like the rank-local conv gradients.
            gflat.mul_(plan.sync_bn_scale(dev, _sync_world()))
        gX = grads.get("X") if ctx.need_x else None
        return _ModelFunction._grads_out(ctx, plan, gX, gflat)


def supported(model):
    """True when every side of the model fits the engine kernels (widths <= 128, <= MAX_OPS operators, and
    the resident weight block Cin x Fout inside the shared-memory budget of both directions:
    ``hgnn_lg_side_fits``).  Otherwise the model runs layer by layer on the per-module kernels."""
    h2 = 2 * model.n_features
    if not (h2 <= 128 and model.featuremap_in[0] <= 128 and model.J + 2 <= _lib.MAX_OPS):
        return False
    plan = get_plan(model)
    if plan.fits is None:
        plan.fits = all(_lib.lib.hgnn_lg_side_fits(plan.K, s.Fs, s.Fc, s.Fout) == 1 for s in plan.sides)
    return plan.fits


def run_model(model, pack, Xp, XLp, xl_is_degree=False):
    """Forward of the whole layer stack on packed rows; returns (bs, dim_output).  ``xl_is_degree``: XLp is the
    pack's own line-graph degree (pack.PackTensor), so the persistent kernels may collapse phantom rows."""
    plan = get_plan(model)
    if model.training and torch.is_grad_enabled():
        # dist.FlatParams (fused_grad): ONE leaf aliasing all parameters takes the flat gradient, instead
        # of ~230 per-parameter views and AccumulateGrad nodes per step
        from .dist import flat_params_of
        fp = flat_params_of(model)
        if fp is not None and fp.fused_grad and fp.matches(plan.params):
            fp.fused_used = True
            return _ModelFunction.apply(model, pack, Xp, XLp, True, xl_is_degree, fp.flat_leaf)
        return _ModelFunction.apply(model, pack, Xp, XLp, False, xl_is_degree, *plan.params)
    if model.training:      # train mode without autograd: batch statistics, running stats updated
        with torch.no_grad():
            return _ModelFunction.apply(model, pack, Xp, XLp, False, xl_is_degree, *plan.params)
    arena = torch.zeros(1, dtype=torch.float64, device=Xp.device)
    return _forward(plan, pack, Xp, XLp, False, arena)[1]
