"""Per-phase host timestamps of the synchronous e2e loop (prepare_batch -> step -> loss.item())."""
import sys, time, gc; sys.path.insert(0, '/root/repo')
import torch, numpy as np
import hgnn_b200
from hgnn_b200 import synth
from hgnn_b200.functions.batching import prepare_batch
from hgnn_b200.models.gnns.model_mnb import GNN_lg
from hgnn_b200.dist import FlatParams, FusedAdamax
inst = synth.sbm_dataset(32, N=1000, sparse=True)
model = GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train(); fp = FlatParams(model); opt = FusedAdamax(fp)
from hgnn_b200 import pack as _pk
_hp, _fg = _pk.host_pack, _pk.BatchPack.from_graphs.__func__
T_HP, T_FG = [0.0, 0], [0.0, 0]
T_C = [0, 0, 0, 0, 0]
def hp(*a, **k):
    t = time.perf_counter(); r = _hp(*a, **k); T_HP[0] += time.perf_counter() - t; T_HP[1] += 1
    
    for q in range(5): T_C[q] += hgnn_b200._lib.lib.hgnn_host_pack_last_ns(q)
    return r
def fg(cls, *a, **k):
    t = time.perf_counter(); r = _fg(cls, *a, **k); T_FG[0] += time.perf_counter() - t; T_FG[1] += 1; return r
_pk.host_pack = hp
_pk.BatchPack.from_graphs = classmethod(fg)
names = ["prepare", "to_device", "zero_grad", "forward", "loss", "backward", "allreduce+opt", "item"]
def step(acc, sync_each):
    ts = [time.perf_counter()]
    def mark():
        if sync_each: torch.cuda.synchronize()
        ts.append(time.perf_counter())
    b = prepare_batch(inst, 0, 1); mark()
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    Xd, XLd, y = X.cuda(non_blocking=True), XL.cuda(non_blocking=True), T.squeeze(1).long().cuda(non_blocking=True); mark()
    fp.zero_grad(); mark()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg); mark()
    loss = torch.nn.functional.cross_entropy(out, y); mark()
    loss.backward(); mark()
    fp.all_reduce_grad(); opt.step(); mark()
    loss.item(); mark()
    for i in range(len(names)): acc[i] += ts[i + 1] - ts[i]
for sync_each in (False, True):
    for _ in range(6): step([0] * 8, sync_each)
    T_HP[:] = [0.0, 0]; T_FG[:] = [0.0, 0]
    acc = [0.0] * 8
    gc0 = gc.get_stats()[2]["collections"]
    t = time.perf_counter()
    for _ in range(20): step(acc, sync_each)
    tot = (time.perf_counter() - t) / 20 * 1e3
    print("host_pack avg %.3f ms, from_graphs avg %.3f ms" % (T_HP[0] / max(T_HP[1], 1) * 1e3, T_FG[0] / max(T_FG[1], 1) * 1e3)); T_HP[:] = [0.0, 0]; T_FG[:] = [0.0, 0]
    print("sync_each=%s total %.3f ms/step  gen2 collections %d" % (sync_each, tot, gc.get_stats()[2]["collections"] - gc0))
    print("   " + " | ".join("%s %.3f" % (n, a / 20 * 1e3) for n, a in zip(names, acc)))
gc.disable()
acc = [0.0] * 8
t = time.perf_counter()
for _ in range(20): step(acc, False)
print("gc disabled total %.3f ms/step" % ((time.perf_counter() - t) / 20 * 1e3))
print("   " + " | ".join("%s %.3f" % (n, a / 20 * 1e3) for n, a in zip(names, acc)))

for nt in (1, 2, 4, 8):
    _pk.HOST_PACK_THREADS = nt
    acc = [0.0] * 8; T_HP[:] = [0.0, 0]; T_FG[:] = [0.0, 0]; T_C[:] = [0, 0, 0, 0, 0]
    t = time.perf_counter()
    for _ in range(20): step(acc, False)
    print("threads %d: total %.3f ms/step, prepare %.3f, host_pack %.3f (C: tasks %.3f, copies %.3f; workers start after %.3f..%.3f ms; caller copied %.1f MB), from_graphs %.3f" % (nt, (time.perf_counter() - t) / 20 * 1e3, acc[0] / 20 * 1e3, T_HP[0] / 20 * 1e3, T_C[0] / 20e6, T_C[1] / 20e6, T_C[3] / 20e6, T_C[2] / 20e6, T_C[4] / 20e6, T_FG[0] / 20 * 1e3)); T_C[:] = [0, 0, 0, 0, 0]
