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
pr=cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
