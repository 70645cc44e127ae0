"""Per-entry-point device times of one real training step (bench.py's profile_step: each recorded launch
re-issued 20x inside a CUDA graph, warm and cold L2) without the rest of the bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import synth  # noqa: E402
from hgnn_b200.dist import FlatParams, FusedAdamax  # noqa: E402
from hgnn_b200.functions.batching import prepare_batch  # noqa: E402
from hgnn_b200.models.gnns.model_mnb import GNN_lg  # noqa: E402

h = int(os.environ.get("KT_H", "2"))
inst = synth.sbm_dataset(32, N=1000, J=1, sparse=True)
torch.manual_seed(0)
model = GNN_lg(0, h, 20, 5, 2, 1, 1).cuda().train()
fp = FlatParams(model)
opt = FusedAdamax(fp)
X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = prepare_batch(inst, 0, 1)
res = (X.cuda(), XL.cuda(), W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, T.squeeze(1).long().cuda())


def train_step(b):
    Xd, XLd, W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, yd = b
    fp.zero_grad()
    out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, yd)
    loss.backward()
    fp.gather_grad()
    opt.step()
    return loss


for _ in range(3):
    train_step(res)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
prof = bench.profile_step(train_step, res, flush)
tot = sum(v["n"] * v["warm_us"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["n"] * kv[1]["warm_us"]):
    if os.environ.get("KT_ALL") != "1" and k[1] not in ("edge", "node"):
        continue
    print("%-24s %-12s n=%2d warm %7.2f us  cold %7.2f us" % (k[0], k[1], v["n"], v["warm_us"], v["cold_us"]))
print("sum over the step (warm): %.1f us" % tot)
