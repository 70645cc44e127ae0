"""Driver for ncu captures of the kernels around the model path (VERDICT r1 "missing" item 7):
pack_gather_kernel (prepare_batch), spgemm_* (J = 2 powers), gmul_kernel (stand-alone graph_oper / P_multi),
ccn2_fwd / ccn2_bwd (CCN_2D on 256 QM9-shaped graphs).
    ncu --set full -k regex:"pack_gather|spgemm|gmul|ccn2" python profiles/prof_misc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import synth  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.functions.utils_ccn import CcnStructure  # noqa: E402
from hgnn_b200.models.compnets.model_ccn import CCN_2D  # noqa: E402
from hgnn_b200.models.layers.layers_mnb import P_multi, graph_oper  # noqa: E402

# ---- prepare_batch (pack_gather_kernel) with J = 2 (spgemm_* for A^2, AL^2 and their transposes)
inst = synth.sbm_dataset(32, N=1000, J=2)
X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 2)
torch.cuda.synchronize()
# ---- stand-alone gmul: graph_oper on W and WL, P_multi on Pm (reference layers_mnb.py:391-434)
gen = torch.Generator().manual_seed(0)
xn = torch.randn(32, 4, W.pack.Nmax, generator=gen).cuda().requires_grad_()
xe = torch.randn(32, 4, W.pack.Emax, generator=gen).cuda().requires_grad_()
y = graph_oper()(W, xn).sum() + graph_oper()(WL, xe).sum() + P_multi()(Pm, xe).sum()
y.backward()
torch.cuda.synchronize()
# ---- CCN-2D, 256 QM9-shaped graphs in one launch group, forward + backward
insts = synth.qm9_shaped_dataset(256)
As = [i[1] + torch.eye(i[1].shape[0]) for i in insts]
st = CcnStructure.from_graphs(As)
Xc = torch.cat([i[0] for i in insts], 0).cuda()
net = CCN_2D(5, 1, 2, 2, True).cuda()
out = net.fc(net._levels(Xc.index_select(0, st.row_vertex2), st))
out.sum().backward()
torch.cuda.synchronize()
print("ok", float(out.sum()), float(y))
