"""Where does the BatchLoader loop spend its time?  Producer: prepare_batch time and time blocked on a full queue;
consumer: time blocked on an empty queue, step issue time, loss.item() wait."""
import sys, time, threading, queue; sys.path.insert(0, '/root/repo')
import torch
import hgnn_b200
from hgnn_b200 import synth
from hgnn_b200.functions import batching
from hgnn_b200.functions.batching import prepare_batch
from hgnn_b200.models.gnns.model_mnb import GNN_lg
from hgnn_b200.dist import FlatParams, FusedAdamax
inst = synth.sbm_dataset(64, N=1000, sparse=True)
model = GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train(); fp = FlatParams(model); opt = FusedAdamax(fp)
P = {"prep": 0.0, "n": 0}
_pb = batching.prepare_batch
def timed_prepare(*a, **k):
    t = time.perf_counter(); r = _pb(*a, **k); P["prep"] += time.perf_counter() - t; P["n"] += 1; return r
batching.prepare_batch = timed_prepare
def step_only(b):
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    Xd, XLd, y = X.cuda(non_blocking=True), XL.cuda(non_blocking=True), T.squeeze(1).long().cuda(non_blocking=True)
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, y); loss.backward(); fp.all_reduce_grad(); opt.step()
    return loss
idx = [list(range(32 * (k % 2), 32 * (k % 2) + 32)) for k in range(46)]
for depth in (1, 2, 3):
    P["prep"], P["n"] = 0.0, 0
    t_get = t_issue = t_item = 0.0
    it = iter(batching.BatchLoader(inst, idx, 0, 1, depth=depth))
    k = 0
    while True:
        t0 = time.perf_counter()
        try: b = next(it)
        except StopIteration: break
        t1 = time.perf_counter()
        loss = step_only(b)
        t2 = time.perf_counter()
        loss.item()
        t3 = time.perf_counter()
        if k == 5: torch.cuda.synchronize(); T0 = time.perf_counter(); t_get = t_issue = t_item = 0.0; P["prep"], P["n"] = 0.0, 0
        elif k > 5: t_get += t1 - t0; t_issue += t2 - t1; t_item += t3 - t2
        k += 1
    n = k - 6
    print("depth %d: %.3f ms/step | consumer: wait-for-batch %.3f, issue %.3f, item %.3f | producer prepare_batch %.3f ms avg"
          % (depth, (time.perf_counter() - T0) / n * 1e3, t_get / n * 1e3, t_issue / n * 1e3, t_item / n * 1e3, P["prep"] / max(P["n"], 1) * 1e3))
# the same without a thread: prepare, then step (sync loop), 2 alternating batches
for _ in range(5): step_only(_pb([inst[i] for i in idx[0]], 0, 1)).item()
torch.cuda.synchronize(); t = time.perf_counter()
for k in range(20): step_only(_pb([inst[i] for i in idx[k]], 0, 1)).item()
print("sync loop: %.3f ms/step" % ((time.perf_counter() - t) / 20 * 1e3))
