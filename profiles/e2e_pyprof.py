"""cProfile of the three host phases of the synchronous e2e loop (C2), each profiled on its own with the device idle
before the phase: prepare_batch, model forward, loss.backward().  Sorted by own time."""
import cProfile
import os
import pstats
import sys

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
prof = {k: cProfile.Profile() for k in ("prepare", "forward", "backward")}
N = 60
for k in range(N + 5):
    on = k >= 5
    torch.cuda.synchronize()
    if on: prof["prepare"].enable()
    b = prepare_batch(hosts[k % 2], 0, 1)
    if on: prof["prepare"].disable()
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    y = T.squeeze(1).long()
    Xd, XLd, yd = X.pin_memory().cuda(non_blocking=True), XL.pin_memory().cuda(non_blocking=True), y.pin_memory().cuda(non_blocking=True)
    fp.zero_grad()
    torch.cuda.synchronize()
    if on: prof["forward"].enable()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    if on: prof["forward"].disable()
    loss = torch.nn.functional.cross_entropy(out, yd)
    torch.cuda.synchronize()
    if on: prof["backward"].enable()
    loss.backward()
    if on: prof["backward"].disable()
    fp.all_reduce_grad(); opt.step()
    loss.item()
for name, pr in prof.items():
    print("=" * 30, name, "(per call = totals / %d)" % N)
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)
