import sys, time; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import hgnn_b200
from hgnn_b200 import synth
from hgnn_b200.functions import batching
from hgnn_b200.functions.batching import prepare_batch
from hgnn_b200.models.gnns.model_mnb import GNN_lg
from hgnn_b200.dist import FlatParams, FusedAdamax
import cProfile, pstats
inst = synth.sbm_dataset(32, N=1000)
model = GNN_lg(0,2,20,5,2,1,1).cuda().train(); fp=FlatParams(model); opt=FusedAdamax(fp)
def step():
    b = prepare_batch(inst, 0, 1)
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    Xd, XLd, y = X.cuda(non_blocking=True), XL.cuda(non_blocking=True), T.squeeze(1).long().cuda(non_blocking=True)
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, y); loss.backward(); fp.all_reduce_grad(); opt.step()
    return loss.item()
for _ in range(3): step()
t=time.perf_counter(); 
for _ in range(10): step()
print('e2e ms', (time.perf_counter()-t)*100)
t=time.perf_counter()
for _ in range(10): b = prepare_batch(inst, 0, 1); torch.cuda.synchronize()
print('prepare_batch ms', (time.perf_counter()-t)*100)
t=time.perf_counter()
for _ in range(10): b = prepare_batch(inst, 0, 1)
print('prepare_batch host-only ms', (time.perf_counter()-t)*100); torch.cuda.synchronize()
from hgnn_b200 import pack as _pk
gs=[i[3].graph_ops for i in inst]
hold={}
t=time.perf_counter()
for _ in range(10): _pk.host_pack(gs, True, True, alloc=_pk._pinned_alloc(hold))
print('host_pack (pinned) ms', (time.perf_counter()-t)*100)
def step_only(b):
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = b
    Xd, XLd, y = X.cuda(non_blocking=True), XL.cuda(non_blocking=True), T.squeeze(1).long().cuda(non_blocking=True)
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, y); loss.backward(); fp.all_reduce_grad(); opt.step()
    return loss
for _ in range(3): step_only(b)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): step_only(b)
th=time.perf_counter()-t; torch.cuda.synchronize()
print('step host issue ms', th*100, ' step incl. device ms', (time.perf_counter()-t)*100)
from hgnn_b200.functions.batching import BatchLoader
t0=None
for k, b in enumerate(BatchLoader(inst, [list(range(32))]*14, 0, 1)):
    if k == 4: torch.cuda.synchronize(); t0=time.perf_counter()
    step_only(b).item()
torch.cuda.synchronize(); print('e2e loader ms', (time.perf_counter()-t0)*100)
pr=cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
