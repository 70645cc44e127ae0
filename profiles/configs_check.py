"""Runs every BASELINE.json config through the public API for a few training steps and prints graphs/s
(eager, host-side batch preparation excluded unless stated).  Not the bench line - a coverage check
that all five configurations execute on the CUDA path at their stated sizes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import synth  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.compnets.model_ccn import CCN_2D  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg, GNN_simple  # noqa: E402


def time_steps(step, n=10, warm=3):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n


def gnn_case(name, inst, model, regression=False):
    t0 = time.perf_counter()
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    Xd, XLd = X.cuda(), XL.cuda()
    y = T.cuda() if regression else T.squeeze(1).long().cuda()
    opt = torch.optim.Adamax(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        out = (model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg) if model.dual
               else model([Xd, W], N_batch, mask))
        loss = torch.nn.functional.mse_loss(out, y) if regression else torch.nn.functional.cross_entropy(out, y)
        loss.backward()
        opt.step()
    dt = time_steps(step)
    print("%-58s bs=%4d  %8.2f ms/step  %10.0f graphs/s   (prepare_batch %.1f ms; rows n=%d m=%d)"
          % (name, len(inst), dt * 1e3, len(inst) / dt, t_prep * 1e3, W.pack.Rn, W.pack.Rm))


torch.manual_seed(0)
gnn_case("C1 GNN SBM N=50 (L=20,h=2,J=1)", synth.sbm_dataset(30, N=50, a=8.0, b=2.0), GNN_simple(0, 2, 20, 5, 2, 1).cuda().train())
gnn_case("C2 LGNN SBM N=1000 (L=20,h=2,J=1,order 1)", synth.sbm_dataset(32, N=1000), GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train())
gnn_case("C3 GNN QM9-shaped (L=15,h=1,J=1), MSE", synth.qm9_shaped_dataset(512), GNN_simple(0, 1, 15, 5, 1, 1).cuda().train(), True)
gnn_case("C3' LGNN QM9-shaped (L=8,h=2, order 2)", synth.qm9_shaped_dataset(512), GNN_lg(0, 2, 8, 5, 1, 1, 2).cuda().train(), True)
t0 = time.perf_counter()
big = synth.sbm_dataset(8, N=10000, a=15.0, b=5.0)
print("C4 dataset build (8 graphs, N=10000, edge-list instances): %.1f s" % (time.perf_counter() - t0))
gnn_case("C4 LGNN SBM N=10000 d~10 (L=20,h=2)", big, GNN_lg(0, 2, 20, 5, 2, 1, 1).cuda().train())

# C5: CCN-2D on 256 QM9-shaped graphs, one launch group
insts = synth.qm9_shaped_dataset(256)
Xs = [i[0].cuda() for i in insts]
As = [i[1] + torch.eye(i[1].shape[0]) for i in insts]
net = CCN_2D(5, 1, 2, 2, True).cuda()
from hgnn_b200.functions.utils_ccn import CcnStructure  # noqa: E402
st = CcnStructure.from_graphs(As)
y = torch.stack([i[2][:1] for i in insts]).cuda()
opt = torch.optim.Adamax(net.parameters(), lr=1e-3)


def ccn_step():
    opt.zero_grad(set_to_none=True)
    out = net.forward_batch(Xs, As, structure=st)
    torch.nn.functional.mse_loss(out, y).backward()
    opt.step()


dt = time_steps(ccn_step)
print("%-58s bs=%4d  %8.2f ms/step  %10.0f graphs/s   (V=%d vertices, nmax=%d)"
      % ("C5 CCN-2D QM9-shaped (layers=2,h=2), batched", 256, dt * 1e3, 256 / dt, st.V, st.nmax))
