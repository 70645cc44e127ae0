"""Per-phase timeline of the persistent kernels (csrc/mega.cu) on the C2 workload.

Every CTA stamps %globaltimer at four points of every phase (phase entered, barrier passed, batch-norm vectors
ready, rows + flush done).  Printed per phase: how long the slowest CTA waited in the barrier, the row work
(median / max over CTAs) and the phase span from the first CTA entering to the last CTA finishing.

    HGNN_B200_MEGA=1 python profiles/mega_trace.py [--bs 32] [--nodes 1000] [--layers 20]
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=32)
    ap.add_argument("--nodes", type=int, default=1000)
    ap.add_argument("--layers", type=int, default=20)
    ap.add_argument("--order", type=int, default=1)
    a = ap.parse_args()
    import hgnn_b200
    from hgnn_b200 import _lib, synth
    from hgnn_b200.functions.batching import prepare_batch
    from hgnn_b200.models.gnns.model_mnb import GNN_lg
    torch.manual_seed(0)
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(synth.sbm_dataset(a.bs, N=a.nodes), 0, 1)
    model = GNN_lg(0, 2, a.layers, 5, 2, 1, a.order).cuda().train()
    Xd, XLd, y = X.cuda(), XL.cuda(), T.squeeze(1).long().cuda()
    pack = W.pack
    grid = int(_lib.lib.hgnn_mega_grid_for(pack.Rn, int(pack.erow.numel())))
    MAXS = 48
    trace = torch.zeros(2 * MAXS * grid * 4, dtype=torch.int64, device="cuda")

    def step():
        out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
        torch.nn.functional.cross_entropy(out, y).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.lib.hgnn_mega_set_trace(ctypes.c_void_p(trace.data_ptr()))
    step()
    torch.cuda.synchronize()
    _lib.lib.hgnn_mega_set_trace(None)
    t = trace.cpu().numpy().reshape(2, MAXS, grid, 4).astype(np.float64)
    print("grid %d CTAs, Rn %d, Rm %d, active line-graph rows %d" % (grid, pack.Rn, pack.Rm, int(pack.erow.numel())))
    for p, name in ((0, "forward"), (1, "backward")):
        ph = [s for s in range(MAXS) if t[p, s, :, 3].max() > 0]
        if not ph:
            continue
        order = ph if p == 0 else ph[::-1]
        t0 = t[p, order[0], :, 0].min()
        tot = (t[p, order[-1], :, 3].max() - t0) / 1e3
        print("%s: %d phases, %.1f us from the first stamp to the last" % (name, len(order), tot))
        print("  phase  enter(first..last)  barrier wait med/max  vectors med  rows+flush med/max  span")
        for s in order:
            e, b, v, r = (t[p, s, :, k] for k in range(4))
            print("  %3d    %7.2f..%7.2f      %6.2f / %6.2f       %6.2f      %6.2f / %6.2f     %6.2f" % (
                s, (e.min() - t0) / 1e3, (e.max() - t0) / 1e3, np.median(b - e) / 1e3, (b - e).max() / 1e3,
                np.median(v - b) / 1e3, np.median(r - v) / 1e3, (r - v).max() / 1e3, (r.max() - e.min()) / 1e3))


if __name__ == "__main__":
    main()
