"""Why is host_pack 4x slower inside the training loop than in a hot loop?  Cold-cache runs into
different kinds of destination memory (torch pinned, pageable numpy, hugepage mmap + cudaHostRegister)."""
import sys, time, mmap, ctypes; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import hgnn_b200
from hgnn_b200 import synth, pack
inst = synth.sbm_dataset(32, N=1000, J=1, sparse=True)
gs = [i[3].graph_ops for i in inst]
n = 24 << 20
torch.cuda.init()
junk = np.empty(256 << 20, np.uint8)
def make(kind):
    if kind == "torch_pinned":
        t = torch.empty(n, dtype=torch.uint8, pin_memory=True); return t, t.numpy()
    if kind == "numpy":
        a = np.empty(n, np.uint8); a[:] = 0; return a, a
    if kind == "huge_registered":
        m = mmap.mmap(-1, n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        a = np.frombuffer(m, dtype=np.uint8)
        off = (-a.ctypes.data) % (2 << 20)
        a = a[off:off + n]
        libc = ctypes.CDLL("libc.so.6")
        print("madvise rc", libc.madvise(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(n), 14))
        a[:] = 0
        rc = torch.cuda.cudart().cudaHostRegister(a.ctypes.data, n, 0)
        print("cudaHostRegister rc", rc)
        return (m, a), a
for kind in ("torch_pinned", "numpy", "huge_registered"):
    bufs = [make(kind) for _ in range(4)]
    for mode in ("hot", "cold"):
        for nt in (1, 8):
            tot = 0.0
            for k in range(12):
                dst = bufs[k % 4][1] if mode == "cold" else bufs[0][1]
                if mode == "cold":
                    junk[:] = k
                    t = time.perf_counter()
                    while time.perf_counter() - t < 0.002: pass
                t = time.perf_counter()
                pack.host_pack(gs, True, True, alloc=lambda nb: dst[:nb], n_threads=nt)
                if k >= 2: tot += time.perf_counter() - t
            print("%-16s %-4s threads %d: %.3f ms" % (kind, mode, nt, tot / 10 * 1e3))
# H2D from the registered hugepage buffer
a = bufs[0][1]
d = torch.empty(n, dtype=torch.uint8, device="cuda")
src = torch.from_numpy(a)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(src, non_blocking=True); e1.record(); e1.synchronize()
print("H2D from registered hugepage numpy buffer: %.3f ms (is_pinned=%s)" % (e0.elapsed_time(e1), src.is_pinned()))
