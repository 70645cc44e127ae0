"""Per-CTA timeline of the width-4 backward kernels (HGNN_B200_ABLATE=8 makes every CTA record its start / end
globaltimer): which CTAs are the slow ones?  Runs the per-side Python path so that each launch can be read back."""
import os
import sys

os.environ["HGNN_B200_ABLATE"] = "8"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import hgnn_b200  # noqa: E402
from hgnn_b200 import _lib, engine, synth  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402

engine.USE_PROGRAM = False
inst = synth.sbm_dataset(32, N=1000)
X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
model = GNN_lg(0, 2, 6, 5, 2, 1, 1).cuda().train()
Xd, XLd, y = X.cuda(), XL.cuda(), T.squeeze(1).long().cuda()
orig = _lib.call
seen = {}


def rec(name, *args):
    rc = orig(name, *args)
    if name == "hgnn_lg_side_bwd" and (_lib.tag.endswith(".edge") or _lib.tag.endswith(".node")) and not _lib.tag.startswith("L0."):
        kind = _lib.tag.split(".")[-1]
        if kind not in seen or seen[kind][0] < 2:          # keep the third launch of each kind (warm)
            buf = (ctypes.c_ulonglong * (3 * 592))()
            orig("hgnn_debug_cta_times", buf, 592)
            ph = (ctypes.c_ulonglong * (3 * 592))()
            orig("hgnn_debug_cta_phases", ph, 592)
            seen[kind] = (seen.get(kind, (0, None))[0] + 1, np.array(buf, dtype=np.int64).reshape(-1, 3),
                          np.array(ph, dtype=np.int64).reshape(-1, 3))
    return rc


_lib.call = engine.call = rec
for _ in range(2):
    for p in model.parameters():
        p.grad = None
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    torch.nn.functional.cross_entropy(out, y).backward()
torch.cuda.synchronize()
for kind, (_, t, ph) in seen.items():
    t0 = t[:, 0].min()
    start, end, role = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2]
    dur = end - start
    print("== bwd %s side: kernel span %.1f us; CTA start spread %.1f us" % (kind, end.max(), start.max()))
    for r, name in ((1, "self"), (0, "cross")):
        m = role == r
        if m.any():
            print("   %-5s CTAs %3d: duration min %.1f  median %.1f  p90 %.1f  max %.1f us; last end %.1f us"
                  % (name, m.sum(), dur[m].min(), np.median(dur[m]), np.percentile(dur[m], 90), dur[m].max(), end[m].max()))
    selfs = np.nonzero(role == 1)[0]
    loop_end, rng_end, nfl = (ph[:, 0] - t0) / 1e3, (ph[:, 1] - t0) / 1e3, ph[:, 2]
    fl = selfs[nfl[selfs] > 0]
    nofl = selfs[nfl[selfs] == 0]
    if fl.size:
        print("   self CTAs with flagged rows %d: row loop ends median %.1f, range phase takes median %.1f max %.1f us, "
              "kernel end median %.1f; without flagged rows %d: row loop ends median %.1f max %.1f, end median %.1f max %.1f"
              % (fl.size, np.median(loop_end[fl]), np.median(rng_end[fl] - loop_end[fl]), (rng_end[fl] - loop_end[fl]).max(),
                 np.median(end[fl]), nofl.size, np.median(loop_end[nofl]), loop_end[nofl].max(), np.median(end[nofl]), end[nofl].max()))
    slow = np.argsort(-end)[:8]
    print("   slowest: (index, flagged rows, row loop end, range phase end, end):",
          [(int(i), int(nfl[i]), round(float(loop_end[i]), 1), round(float(rng_end[i]), 1), round(float(end[i]), 1)) for i in slow if role[i] == 1])
    print("   slowest CTAs (index, role, start, end):", [(int(i), "self" if role[i] else "cross", round(float(start[i]), 1), round(float(end[i]), 1)) for i in slow])
