import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import hgnn_b200
from hgnn_b200 import synth
from hgnn_b200.dist import FlatParams, FusedAdamax
from hgnn_b200.functions.batching import prepare_batch
from hgnn_b200.models.gnns.model_mnb import GNN_lg
hosts = [synth.sbm_dataset(32, N=1000, sparse=True, first_id=k * 32) for k in range(2)]
model = GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train()
fp = FlatParams(model); opt = FusedAdamax(fp)
def dev(b):
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    y = T.squeeze(1).long()
    return (X.pin_memory().cuda(non_blocking=True), XL.pin_memory().cuda(non_blocking=True), W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, y.pin_memory().cuda(non_blocking=True))
def step(d):
    Xd, XLd, W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, yd = d
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, yd)
    loss.backward(); fp.all_reduce_grad(); opt.step()
    return loss
res = dev(prepare_batch(hosts[0], 0, 1))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step(res)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    sl = step(res)
for mode in ("eager", "graph-replay (upper bound: issue cost ~0)"):
    for k in range(6):
        d = dev(prepare_batch(hosts[k % 2], 0, 1))
        (step(d) if mode == "eager" else (g.replay(), sl)[1]).item()
    torch.cuda.synchronize(); t = time.perf_counter()
    for k in range(50):
        d = dev(prepare_batch(hosts[k % 2], 0, 1))
        (step(d) if mode == "eager" else (g.replay(), sl)[1]).item()
    torch.cuda.synchronize()
    print(mode, "%.3f ms/step" % ((time.perf_counter() - t) / 50 * 1e3))
