"""Small driver for ncu captures: one LGNN training step (C2 batch: 32 x SBM N=1000) with few layers,
so `ncu -k regex:... -c N` sees each kernel of the hot path a handful of times.
Usage: python profiles/prof_step.py [--h 2] [--layers 4] [--steps 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=2)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--bs", type=int, default=32)
ap.add_argument("--nodes", type=int, default=1000)
a = ap.parse_args()

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import synth  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402

inst = synth.sbm_dataset(a.bs, N=a.nodes)
X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
model = GNN_lg(0, a.h, a.layers, 5, 2, 1, 1).cuda().train()
Xd, XLd, y = X.cuda(), XL.cuda(), T.squeeze(1).long().cuda()
for _ in range(a.steps):
    for p in model.parameters():
        p.grad = None
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    torch.nn.functional.cross_entropy(out, y).backward()
torch.cuda.synchronize()
print("ok", float(out.sum()))
