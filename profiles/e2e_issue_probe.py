"""Where the host time of a pass goes (C2, synchronous e2e loop): wall time of the two native executor calls
(hgnn_program_fwd / _bwd, ~40 launches each) against the Python around them, with the launch queue empty at the
start of the pass (device idle: pure issue cost) and in the normal loop."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import _lib, synth  # noqa: E402
from hgnn_b200.dist import FlatParams, FusedAdamax  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402

hosts = [synth.sbm_dataset(32, N=1000, sparse=True, first_id=k * 32) for k in range(2)]
model = GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train()
fp = FlatParams(model)
opt = FusedAdamax(fp)

native = {}
_orig = _lib.call_program


def timed_call_program(name, *args):
    t = time.perf_counter()
    _orig(name, *args)
    native[name] = native.get(name, 0.0) + time.perf_counter() - t


import hgnn_b200.engine as engine  # noqa: E402
engine.call_program = timed_call_program

names = ["prepare", "to_device", "forward", "loss", "backward", "opt", "item"]


def step(k, acc, drain):
    ts = [time.perf_counter()]
    b = prepare_batch(hosts[k % 2], 0, 1); ts.append(time.perf_counter())
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    y = T.squeeze(1).long()
    Xd, XLd, yd = X.pin_memory().cuda(non_blocking=True), XL.pin_memory().cuda(non_blocking=True), y.pin_memory().cuda(non_blocking=True)
    if drain:
        torch.cuda.synchronize()
    ts.append(time.perf_counter())
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg); ts.append(time.perf_counter())
    loss = torch.nn.functional.cross_entropy(out, yd); ts.append(time.perf_counter())
    loss.backward(); ts.append(time.perf_counter())
    fp.all_reduce_grad(); opt.step(); ts.append(time.perf_counter())
    loss.item(); ts.append(time.perf_counter())
    if acc is not None:
        for i in range(len(names)):
            acc[i] += ts[i + 1] - ts[i]


for drain in (False, True):
    for k in range(8):
        step(k, None, drain)
    native.clear()
    acc = [0.0] * len(names)
    n = 40
    t = time.perf_counter()
    for k in range(n):
        step(k, acc, drain)
    print("drain=%s total %.3f ms/step" % (drain, (time.perf_counter() - t) / n * 1e3))
    print("   " + " | ".join("%s %.3f" % (nm, a / n * 1e3) for nm, a in zip(names, acc)))
    print("   native: " + " | ".join("%s %.3f" % (k_, v / n * 1e3) for k_, v in sorted(native.items())))
