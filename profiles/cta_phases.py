"""Per-CTA phase breakdown of the width-4 forward and backward kernels (HGNN_B200_ABLATE=8): start -> coefficient vectors in shared
memory -> end of the row loop -> end (flush done), medians / p90 / max over the self and the cross CTAs.  Per-side Python
path (eager launches, no PDL overlap)."""
import os
import sys

os.environ["HGNN_B200_ABLATE"] = "8"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
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
NC = 1024


def rec(name, *args):
    rc = orig(name, *args)
    if name in ("hgnn_lg_side_bwd", "hgnn_lg_side_fwd") and (_lib.tag.endswith(".edge") or _lib.tag.endswith(".node")) \
            and not _lib.tag.startswith("L0.") and not _lib.tag.startswith("layer0"):
        kind = ("bwd " if name.endswith("bwd") else "fwd ") + _lib.tag.split(".")[-1]
        buf = (ctypes.c_ulonglong * (3 * NC))()
        orig("hgnn_debug_cta_times", buf, NC)
        ph = (ctypes.c_ulonglong * (3 * NC))()
        orig("hgnn_debug_cta_phases", ph, NC)
        seen[kind] = (np.array(buf, dtype=np.int64).reshape(-1, 3), np.array(ph, dtype=np.int64).reshape(-1, 3))
    return rc


_lib.call = engine.call = rec
for _ in range(3):
    for p in model.parameters():
        p.grad = None
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    torch.nn.functional.cross_entropy(out, y).backward()
torch.cuda.synchronize()


def q(v):
    return "median %.2f p90 %.2f max %.2f" % (np.median(v), np.percentile(v, 90), v.max())


for kind, (t, ph) in seen.items():
    live = (t[:, 0] > 0) & (t[:, 0] > t[:, 0].max() - 50000)      # slots of this launch only (older launches had more CTAs)
    live &= (t[:, 2] == 2) if kind.startswith("fwd") else (t[:, 2] != 2)
    t, ph = t[live], ph[live]
    t0 = t[:, 0].min()
    start, end, role = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2]
    coef, loop = (ph[:, 2] - t0) / 1e3, (ph[:, 0] - t0) / 1e3
    print("== %s side: %d CTAs, kernel span %.2f us; CTA start spread %.2f us" % (kind, len(t), end.max(), start.max()))
    if kind.startswith("fwd"):
        wait = (ph[:, 1] - t0) / 1e3
        print("   all   CTAs %3d | start->wait passed %s | ->coef %s | coef->rows done %s | rows done->end %s | end at %s"
              % (len(t), q(wait - start), q(coef - wait), q(loop - coef), q(end - loop), q(end)))
        continue
    for r, name in ((1, "self"), (0, "cross")):
        m = (role == r) & (ph[:, 2] > 0)
        if not m.any():
            continue
        print("   %-5s CTAs %3d | start->coef %s | coef->rows done %s | rows done->end %s | end at %s"
              % (name, m.sum(), q(coef[m] - start[m]), q(loop[m] - coef[m]), q(end[m] - loop[m]), q(end[m])))
