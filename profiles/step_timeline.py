"""Timeline of the width-4 side launches inside ONE replay of the captured C2 training step (HGNN_B200_ABLATE=16: every
traced kernel records min CTA start / min, max "producer wait passed" / max CTA end with %globaltimer into its slot).
Shows what the chain of 72 dependent launches really costs under PDL inside the graph: per launch the gap to the
previous launch's end, the time before / after the wait, and the totals per kind."""
import ctypes
import os
import sys

os.environ["HGNN_B200_ABLATE"] = str(16 | int(os.environ.get("HGNN_B200_ABLATE", "0")))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import _lib, synth  # noqa: E402
from hgnn_b200.dist import FlatParams, FusedAdamax  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402

L = int(os.environ.get("LAYERS", "20"))
inst = synth.sbm_dataset(32, N=1000, sparse=True)
X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
model = GNN_lg(0, 2, L, 5, 2, 1, 1).cuda().train()
fp = FlatParams(model)
opt = FusedAdamax(fp)
Xd, XLd, y = X.cuda(), XL.cuda(), T.squeeze(1).long().cuda()


def train_step():
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, y)
    loss.backward()
    fp.all_reduce_grad()
    opt.step()
    return loss


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        train_step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
_lib.call("hgnn_debug_ktrace", None, 0, 1)           # numbering restarts: the captured launches take slots 0..
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    train_step()
n = 2 * 2 * (L - 2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
rows = []
for rep in range(5):
    flush.zero_()
    _lib.call("hgnn_debug_ktrace", None, 0, 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (4 * n))()
    _lib.call("hgnn_debug_ktrace", buf, n, 0)
    rows.append((e0.elapsed_time(e1), np.array(buf, dtype=np.int64).reshape(n, 4)))
ms, t = sorted(rows, key=lambda r: r[0])[len(rows) // 2]
print("replay %.3f ms (median of 5, L2 flushed before); %d traced launches" % (ms, n))
half = n // 2
kinds = ["fwd node" if k % 2 == 0 else "fwd edge" for k in range(half)] + ["bwd edge" if k % 2 == 0 else "bwd node" for k in range(half)]
t0 = t[0, 0]
print("%-4s %-9s %9s %9s %9s %9s %9s %9s" % ("slot", "kind", "start", "gap", "pre-wait", "wait-skew", "post-wait", "span"))
agg = {}
for k in range(n):
    st, w0, w1, en = t[k]
    gap = (st - t[k - 1, 3]) / 1e3 if k else 0.0
    rel_gap = (w0 - t[k - 1, 3]) / 1e3 if k else 0.0      # previous kernel fully done -> first CTA past the wait
    row = ((st - t0) / 1e3, gap, (w0 - st) / 1e3, (w1 - w0) / 1e3, (en - w1) / 1e3, (en - st) / 1e3, rel_gap,
           (en - t[k - 1, 3]) / 1e3 if k else 0.0)
    if k < 6 or half - 2 <= k < half + 6 or k >= n - 2:
        print("%-4d %-9s %9.2f %9.2f %9.2f %9.2f %9.2f %9.2f" % ((k, kinds[k]) + row[:6]))
    if 2 <= k < half - 1 or half + 2 <= k:
        agg.setdefault(kinds[k], []).append(row)
print("\nper kind (median over the middle launches, us): gap = start - previous end (negative: started under the producer's tail);")
print("release = previous end -> first CTA past the wait; step = previous end -> this end (what the launch adds to the chain)")
print("%-9s %5s %8s %9s %10s %10s %8s %8s %8s" % ("kind", "n", "gap", "pre-wait", "wait-skew", "post-wait", "span", "release", "step"))
tot = 0.0
for kind, rws in agg.items():
    a = np.array(rws)
    med = np.median(a, axis=0)
    print("%-9s %5d %8.2f %9.2f %10.2f %10.2f %8.2f %8.2f %8.2f" % (kind, len(rws), med[1], med[2], med[3], med[4], med[5], med[6], med[7]))
    tot += a[:, 7].sum()
print("sum of 'step' over the aggregated launches: %.1f us; first traced start -> last traced end: %.1f us"
      % (tot, (t[-1, 3] - t0) / 1e3))
