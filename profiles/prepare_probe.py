"""Phase timing of prepare_batch / BatchPack.from_graphs on the GPU box (host clock, device synchronised between phases)."""
import sys, time, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import hgnn_b200
from hgnn_b200 import synth, pack, _lib
from hgnn_b200.functions.batching import prepare_batch
inst = synth.sbm_dataset(32, N=1000, J=1, sparse=True)
gs = [i[3].graph_ops for i in inst]
dev = torch.device("cuda", 0)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    t0 = T()
    X = torch.zeros(32, 5, 1000, pin_memory=True); XL = torch.zeros(32, 1, 5100, pin_memory=True)
    t1 = T()
    hold = {}
    _, layout = pack.host_pack(gs, True, True, alloc=pack._pinned_alloc(hold))
    t2 = T()
    d = hold["host"].to(dev, non_blocking=True)
    t3 = T()
    views, _ = pack._device_views(hold["host"], layout, dev)
    t4 = T()
    _lib.call("hgnn_fixup_offsets", d.data_ptr(), _lib.iptr(views["fixup"]), views["fixup"].numel() // 4, 32, _lib.stream())
    t5 = T()
    p = pack.BatchPack.from_graphs(gs, 1, True, dev)
    t6 = T()
    b = prepare_batch(inst, 0, 1)
    t7 = T()
    print("it%d  X/XL pinned zeros %.3f | host_pack+alloc %.3f | H2D %.3f | H2D+views %.3f | fixup %.3f | from_graphs %.3f | prepare_batch %.3f ms"
          % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3, (t7-t6)*1e3))
    del hold, d, views, p, b
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    b = prepare_batch(inst, 0, 1); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
