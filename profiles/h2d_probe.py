"""Host-side probe for the e2e path: pinned allocation cost, H2D bandwidth, host_pack thread scaling."""
import sys, time, os; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import hgnn_b200
from hgnn_b200 import synth, pack
print("cpus", os.cpu_count(), "threads default", pack.HOST_PACK_THREADS)
n = 17_562_832
t = time.perf_counter(); h = torch.empty(n, dtype=torch.uint8, pin_memory=True); print("first pinned alloc ms", (time.perf_counter() - t) * 1e3)
del h
t = time.perf_counter(); h = torch.empty(n, dtype=torch.uint8, pin_memory=True); print("second pinned alloc ms", (time.perf_counter() - t) * 1e3)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for size in (n, 1 << 20, 4 << 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d[:size].copy_(h[:size], non_blocking=True); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): d[:size].copy_(h[:size], non_blocking=True)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("H2D %d bytes: %.3f ms = %.1f GB/s" % (size, ms, size / ms / 1e6))
src = np.random.randint(0, 255, n, dtype=np.uint8)
hn = h.numpy()
for _ in range(2):
    t = time.perf_counter(); np.copyto(hn, src); dt = time.perf_counter() - t
print("memcpy pageable->pinned 1 thread: %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
dst = np.empty(n, dtype=np.uint8); np.copyto(dst, src)
t = time.perf_counter(); np.copyto(dst, src); dt = time.perf_counter() - t
print("memcpy pageable->pageable 1 thread: %.2f ms = %.1f GB/s" % (dt * 1e3, n / dt / 1e9))
inst = synth.sbm_dataset(32, N=1000, J=1, sparse=True)
gs = [i[3].graph_ops for i in inst]
for nt in (1, 2, 4, 8, 16):
    pack.host_pack(gs, True, True, alloc=lambda nb: hn[:nb], n_threads=nt)
    t = time.perf_counter()
    for _ in range(10): pack.host_pack(gs, True, True, alloc=lambda nb: hn[:nb], n_threads=nt)
    print("host_pack into pinned, %2d threads: %.3f ms" % (nt, (time.perf_counter() - t) * 100))
hold = {}
t = time.perf_counter()
for _ in range(10): pack.host_pack(gs, True, True, alloc=pack._pinned_alloc(hold), n_threads=8)
print("host_pack + fresh pinned alloc each time: %.3f ms" % ((time.perf_counter() - t) * 100))
