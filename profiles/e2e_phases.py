"""Per-phase host timestamps of the synchronous e2e loop of bench.py (C2): prepare_batch -> H2D -> forward ->
loss -> backward -> all-reduce + Adamax -> loss.item(), with and without a device sync after every phase, and a
cProfile of the loop (top host functions by cumulative time)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import synth  # noqa: E402
from hgnn_b200.dist import FlatParams, FusedAdamax  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402

hosts = [synth.sbm_dataset(32, N=1000, sparse=True, first_id=k * 32) for k in range(2)]
model = GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train()
fp = FlatParams(model)
opt = FusedAdamax(fp)
names = ["prepare", "to_device", "zero_grad", "forward", "loss", "backward", "allreduce+opt", "item"]


def step(k, acc=None, sync_each=False):
    ts = [time.perf_counter()]

    def mark():
        if sync_each:
            torch.cuda.synchronize()
        ts.append(time.perf_counter())
    b = prepare_batch(hosts[k % 2], 0, 1); mark()
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    y = T.squeeze(1).long()
    Xd, XLd, yd = X.pin_memory().cuda(non_blocking=True), XL.pin_memory().cuda(non_blocking=True), y.pin_memory().cuda(non_blocking=True); mark()
    fp.zero_grad(); mark()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg); mark()
    loss = torch.nn.functional.cross_entropy(out, yd); mark()
    loss.backward(); mark()
    fp.all_reduce_grad(); opt.step(); mark()
    loss.item(); mark()
    if acc is not None:
        for i in range(len(names)):
            acc[i] += ts[i + 1] - ts[i]


for sync_each in (False, True):
    for k in range(8):
        step(k, None, sync_each)
    acc = [0.0] * len(names)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for k in range(40):
        step(k, acc, sync_each)
    torch.cuda.synchronize()
    print("sync_each=%s total %.3f ms/step" % (sync_each, (time.perf_counter() - t) / 40 * 1e3))
    print("   " + " | ".join("%s %.3f" % (n, a / 40 * 1e3) for n, a in zip(names, acc)))

pr = cProfile.Profile()
pr.enable()
for k in range(40):
    step(k)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
