"""Wall time of the pieces of prepare_batch (C2), device idle between calls: instance ops lookup, BatchPack.from_graphs
(device_pack: plan + allocations + upload call + views), host tensors + fill loop + handles."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hgnn_b200  # noqa: E402,F401
from hgnn_b200 import _lib, pack as packmod, synth  # noqa: E402
from hgnn_b200.functions import batching  # noqa: E402

hosts = [synth.sbm_dataset(32, N=1000, sparse=True, first_id=k * 32) for k in range(2)]
acc = {}


def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    return w


packmod.BatchPack.from_graphs = classmethod(timed("from_graphs", packmod.BatchPack.from_graphs.__func__))
packmod.device_pack = timed("device_pack", packmod.device_pack)
_orig_call = _lib.call


def call(name, *a):
    t = time.perf_counter()
    _orig_call(name, *a)
    acc["call:" + name] = acc.get("call:" + name, 0.0) + time.perf_counter() - t


_lib.call = packmod.call = call if hasattr(packmod, "call") else call
packmod._lib.call = call
batching._instance_ops = timed("_instance_ops", batching._instance_ops)
N = 200
for k in range(20):
    batching.prepare_batch(hosts[k % 2], 0, 1)
torch.cuda.synchronize()
acc.clear()
tot = 0.0
for k in range(N):
    torch.cuda.synchronize()
    t = time.perf_counter()
    b = batching.prepare_batch(hosts[k % 2], 0, 1)
    tot += time.perf_counter() - t
print("prepare_batch %.3f ms" % (tot / N * 1e3))
for k_, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("   %-36s %.3f ms" % (k_, v / N * 1e3))
