"""Micro-measurements behind profiles/README.md: (1) the per-node floor of a CUDA-graph replay on this
box, (2) each engine entry point launched back to back (no host gaps: N launches, one sync) with a
warm L2, by wrapping the C-ABI calls of one real training step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import _lib, synth  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402


def timed(fn, n=50):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


# (1) graph floor: 85 dependent trivial kernels
x = torch.zeros(1024, device="cuda")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        for _ in range(85):
            x.add_(1.0)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(85):
        x.add_(1.0)
print("graph of 85 trivial dependent kernels: %.1f us per replay = %.2f us per node" % (timed(g.replay, 20), timed(g.replay, 20) / 85))

# (2) real step, every C-ABI call re-issued back to back
inst = synth.sbm_dataset(32, N=1000)
X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
model = GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train()
Xd, XLd, y = X.cuda(), XL.cuda(), T.squeeze(1).long().cuda()
calls = []
orig = _lib.call
hgnn_b200.engine.USE_PROGRAM = False     # the per-side Python loop: one visible C-ABI call per launch
only_sides = os.environ.get("MICROBENCH_ONLY_SIDES") == "1"


def rec(name, *args):
    calls.append((name, _lib.tag, args))
    return orig(name, *args)


for it in range(2):
    calls.clear()
    for p in model.parameters():
        p.grad = None
    _lib.call = rec
    hgnn_b200.engine.call = rec
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, y)
    loss.backward()
    _lib.call = orig
    hgnn_b200.engine.call = orig
torch.cuda.synchronize()
# NOTE: re-issuing a call repeats its atomics into the accumulators: numerically meaningless, same work
seen = {}
for name, tag, args in calls:
    kind = "edge" if tag.endswith(".edge") else "node" if tag.endswith(".node") else tag
    if tag.startswith("L0."):
        kind = "L0." + kind
    key = (name, kind if name.startswith("hgnn_lg_side") else "")
    if key in seen or (only_sides and (not name.startswith("hgnn_lg_side") or kind.startswith("L0") or kind == "readout")):
        continue
    fn = getattr(_lib.lib, name)
    seen[key] = timed(lambda: fn(*args), 40)
for k, v in sorted(seen.items(), key=lambda kv: -kv[1]):
    print("%-32s %-12s %8.2f us back-to-back (warm L2)" % (k[0], k[1], v))
