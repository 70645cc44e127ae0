"""Line-level wall time of prepare_batch (a copy of its body with timers): C2 batch, or `qm9`: 512 QM9-shaped graphs."""
import os, sys, time, ctypes
sys.path.insert(0, "/root/repo" if os.path.isdir("/root/repo/hgnn-2_b200") else os.getcwd())
import numpy as np, torch
import hgnn_b200
from hgnn_b200 import _lib, synth
from hgnn_b200.functions import batching as B
from hgnn_b200.pack import BatchPack, GraphHandle, MaskHandle, OperatorHandle, PackTensor
if len(sys.argv) > 1 and sys.argv[1] == "qm9":
    hosts = [synth.qm9_shaped_dataset(512, J=1, sparse=True, first_id=k * 512) for k in range(2)]
else:
    hosts = [synth.sbm_dataset(32, N=1000, sparse=True, first_id=k * 32) for k in range(2)]
acc = {}
def run(batch, task=0, J=1):
    ts=[time.perf_counter()]
    def mark(): ts.append(time.perf_counter())
    bs = len(batch)
    graphs = [B._instance_ops(inst) for inst in batch]; mark()
    pack = BatchPack.from_graphs(graphs, J, dual=True, device="cuda"); mark()
    n_feat = batch[0][0].shape[1]
    N_batch = torch.tensor([g.N for g in graphs], dtype=torch.int64)
    E_batch = torch.tensor([g.M for g in graphs], dtype=torch.int64)
    Nmax, Emax = int(N_batch.max()), int(E_batch.max()); mark()
    X = torch.empty(bs, n_feat, Nmax, pin_memory=True)
    XL = torch.empty(bs, 1, Emax, pin_memory=True); mark()
    tsl = [inst[2] for inst in batch]
    if all(torch.is_tensor(t) and t.dim() == 1 and t.shape == tsl[0].shape and t.dtype == tsl[0].dtype for t in tsl):
        T = torch.stack(tsl, 0)[:, task].to(torch.float32).reshape(bs, 1)
    mark()
    xs = [inst[0] for inst in batch]
    ok = all(torch.is_tensor(x) and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == n_feat and x.is_contiguous() and not x.is_cuda for x in xs); mark()
    blobs = (ctypes.c_void_p * bs)(*[g.blob_ptr() for g in graphs])
    rows = (ctypes.c_void_p * bs)(*[x.data_ptr() for x in xs]); mark()
    rc = _lib.lib.hgnn_host_fill_features(bs, blobs, rows, n_feat, Nmax, X.data_ptr(), Emax, XL.data_ptr()); mark()
    XL = PackTensor.wrap(XL, pack)
    W, WL = OperatorHandle(pack, "W"), OperatorHandle(pack, "WL")
    Pm, Pd = OperatorHandle(pack, "Pm"), OperatorHandle(pack, "Pd")
    mask, mask_lg = MaskHandle(pack, False), MaskHandle(pack, True); mark()
    names=["instance_ops","from_graphs","N/E tensors","pinned empty x2","T stack","xs check","ctypes arrays","fill call","handles"]
    for i,n in enumerate(names): acc[n]=acc.get(n,0)+ts[i+1]-ts[i]
for k in range(20): run(hosts[k%2])
acc.clear()
N=200
for k in range(N):
    torch.cuda.synchronize(); run(hosts[k%2])
for n,v in acc.items(): print("%-18s %.3f ms"%(n, v/N*1e3))
