#!/usr/bin/env python
"""bench.py -- training throughput of the hot path on the synthetic workloads of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5]

Default (``--config c2`` = BASELINE.json configs[1], the configuration the headline metric is quoted on): one
"step" = one LGNN training step (forward + cross-entropy + backward + gradient all-reduce + Adamax) over one batch
of 32 binary-SBM graphs with N = 1000 nodes per GPU (weak scaling: every rank owns its own 32 graphs; the only
collective is the flat-gradient all-reduce).  Model = ``GNN_lg(task=0, h=2, L=20, dim_input=5, dim_output=2,
J=1, order=1)`` - the reference's script defaults (scripts/main_gnn.py:59,75-77), fp32.

Other configurations (parity-test cases of BASELINE.json; same JSON contract, their own metric names):
  c1  GNN_simple on SBM N = 50, batch 30 (scripts/main_gnn.py defaults; the reference's CPU-runnable case)
  c3  GNN_simple(0, 1, 15, 5, 1, 1) regression on QM9-shaped graphs, batch 512 (scripts/main_gnn_qm9.py:77-78)
  c4  GNN_lg on SBM N = 10 000 (a, b = 15, 5), 8 graphs split over the ranks (strong scaling)
  c5  CCN_2D(5, 1, 2, 2) on QM9-shaped graphs, 256 graphs per step (the reference steps per graph,
      scripts/train_ccn.py:31; batching is the documented departure)

Printed JSON (one line, rank 0):
  value   graphs/s with the batch already packed in HBM (CUDA-graph replay of the whole step, CUDA events per step,
          L2 flushed between steps, max over ranks);
  e2e     graphs/s through the public API from HOST instances: prepare_batch (per-graph DMA out of the pinned
          dataset + GPU gather into block-diagonal CSR) -> model -> loss -> backward -> optimizer -> loss.item();
          e2e.prefetch = the same loop fed by functions.batching.BatchLoader;
  roofline  the dominant aggregation kernel timed alone with CUDA events, algorithmic bytes per SURVEY.md 8(d) /
          DESIGN.md, against MEASURED_PEAKS.json;
  parity  the CUDA model against the CPU oracle on the SAME sample the cpu_baseline leg builds (output and every
          gradient, rel. error; the run fails above 1e-4);
  cpu_baseline  the CPU oracle port (oracle/hgnn_oracle.py = the reference's dense torch.mm loops) timed on the host
          cores on a bounded sample of the same workload; gpu_dense_baseline = the same dense code with .cuda();
  ranks_agree / allreduce_exposed_us  (N > 1) parameter checksums equal on all ranks after the timed steps; step time
          with minus without the gradient all-reduce.
``--impl reference`` times the CPU port alone (the reference is pure Python and cannot travel to the GPU box; see
DESIGN.md); it imports nothing but ``oracle/`` (no CUDA library is loaded).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "graphs/s"
TOL = 1e-4

CONFIGS = {
    "c1": dict(kind="gnn", data="sbm", nodes=50, sbm_a=8.0, sbm_b=2.0, bs=30, layers=20, h=2, J=1, order=0, dim_out=2,
               loss="ce", scaling="weak", metric="gnn_sbm50_train_graphs_per_s", cpu_sample=30,
               what="GNN (GNN_simple, L=%(layers)d, h=%(h)d, J=%(J)d) on 2-class binary SBM N=%(nodes)d (a=8,b=2)"),
    "c2": dict(kind="lgnn", data="sbm", nodes=1000, sbm_a=7.0, sbm_b=3.0, bs=32, layers=20, h=2, J=1, order=1, dim_out=2,
               loss="ce", scaling="weak", metric="lgnn_sbm_train_graphs_per_s", cpu_sample=2,
               what="LGNN (GNN_lg order %(order)d, L=%(layers)d, h=%(h)d, J=%(J)d) on 2-class binary SBM N=%(nodes)d (a=7,b=3)"),
    "c3": dict(kind="gnn", data="qm9", nodes=0, sbm_a=0.0, sbm_b=0.0, bs=512, layers=15, h=1, J=1, order=0, dim_out=1,
               loss="mse", scaling="weak", metric="gnn_qm9_train_graphs_per_s", cpu_sample=512,
               what="GNN regression (GNN_simple, L=%(layers)d, h=%(h)d, J=%(J)d) on synthetic QM9-shaped molecular graphs (<= 29 atoms)"),
    "c4": dict(kind="lgnn", data="sbm", nodes=10000, sbm_a=15.0, sbm_b=5.0, bs=8, layers=20, h=2, J=1, order=1, dim_out=2,
               loss="ce", scaling="strong", metric="lgnn_sbm10k_train_graphs_per_s", cpu_sample=0,
               what="LGNN (GNN_lg order %(order)d, L=%(layers)d, h=%(h)d) on binary SBM N=%(nodes)d (a=15,b=5), 8 graphs over the ranks"),
    "c5": dict(kind="ccn2", data="qm9", nodes=0, sbm_a=0.0, sbm_b=0.0, bs=256, layers=2, h=2, J=1, order=0, dim_out=1,
               loss="mse", scaling="weak", metric="ccn2_qm9_train_graphs_per_s", cpu_sample=16,
               what="CCN-2D (CCN_2D(5, 1, %(h)d, %(layers)d)) second-order covariant contraction on synthetic QM9-shaped graphs"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--bs", type=int, default=None, help="graphs per GPU (c4: graphs in total)")
    ap.add_argument("--nodes", type=int, default=None)
    ap.add_argument("--h", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--order", type=int, default=None)
    ap.add_argument("--J", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline / parity / gpu_dense_baseline legs")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=None, help="graphs per CPU-baseline step")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    for k in ("bs", "nodes", "h", "layers", "order", "J"):
        if getattr(a, k) is None:
            setattr(a, k, cfg[k])
    if a.cpu_sample is None:
        a.cpu_sample = cfg["cpu_sample"]
    for k in ("kind", "data", "sbm_a", "sbm_b", "dim_out", "loss", "scaling", "metric", "what"):
        setattr(a, k, cfg[k])
    return a


def graphs_per_rank(a, world):
    """Weak scaling: a.bs graphs per GPU.  c4 (strong): a.bs graphs split over the ranks."""
    return a.bs if a.scaling == "weak" else max(1, a.bs // max(world, 1))


def workload_config(a, world=None, graphs_per_step=None):
    world = a.gpus if world is None else world
    per = graphs_per_rank(a, world)
    cfg = {"workload": (a.what % vars(a)) + ", batch %d graphs per GPU, fp32 fwd+bwd+Adamax" % per,
           "name": a.config, "graphs_per_gpu": per, "layers": a.layers, "h": a.h, "J": a.J, "order": a.order,
           "parallelism": "dp%d (graphs sharded, flat-gradient all-reduce)" % world,
           "l2": "256 MiB buffer written between timed steps (L2 flush)"}
    if a.data == "sbm":
        cfg["nodes_per_graph"] = a.nodes
    if graphs_per_step is not None:
        cfg["graphs_per_step"] = graphs_per_step
    return cfg


# --------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's dense path, on the host cores.  Imports oracle/ only.
# --------------------------------------------------------------------------------------------
def oracle_instances(a, n_graphs, first_id=0):
    from oracle import hgnn_oracle as O
    from oracle import workloads
    inst = []
    for i in range(n_graphs):
        if a.data == "sbm":
            x, A, t = workloads.sbm_dense(first_id + i, N=a.nodes, a=a.sbm_a, b=a.sbm_b)
        else:
            x, A, t = workloads.qm9_shaped_dense(first_id + i)
        inst.append([x, A, t])
    return O, inst


def oracle_setup(a, n_graphs, first_id=0, device="cpu"):
    """(step function, state) of the dense reference path on n_graphs of the workload; parameters seeded (seed 0)."""
    O, inst = oracle_instances(a, n_graphs, first_id)
    dev = torch.device(device)
    if a.kind == "ccn2":
        p = {k: v.to(dev).requires_grad_() for k, v in O.init_ccn_params(2, 5, a.dim_out, a.h, a.layers, seed=0).items()}
        graphs = [(x.to(dev), (A + torch.eye(A.shape[0])).to(dev), t[:1].to(dev)) for x, A, t in inst]
        opt = torch.optim.Adamax(list(p.values()), lr=1e-3)

        def forward():
            return torch.stack([O.ccn2_forward(p, a.layers, x, A) for x, A, _ in graphs]), torch.stack([t for _, _, t in graphs])
        return O, p, opt, forward, {"n": n_graphs, "dense_mb": 0}
    full = [i + list(O.graph_operators([i[0], i[1]], a.J, True)) for i in inst]
    batch = [t.to(dev) for t in O.prepare_batch(full, 0, a.J)]
    kind = "lg" if a.kind == "lgnn" else "simple"
    p = {k: v.to(dev).requires_grad_() for k, v in
         O.init_gnn_params(kind, a.h, a.layers, 5, a.dim_out, a.J, max(a.order, 1), seed=0).items()}
    opt = torch.optim.Adamax(list(p.values()), lr=1e-3)
    X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = batch
    target = T if a.loss == "mse" else T.squeeze(1).long()

    def forward():
        if a.kind == "lgnn":
            out = O.gnn_lg_forward(p, a.layers, a.order, [X, XL, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
        else:
            out = O.gnn_simple_forward(p, a.layers, [X, W], N_batch, mask)
        return out, target
    return O, p, opt, forward, {"n": n_graphs, "dense_mb": int(WL[0].numel() * 4 / 1e6)}


def loss_of(a, out, target):
    if a.loss == "mse":
        return torch.nn.functional.mse_loss(out, target)
    return torch.nn.functional.cross_entropy(out, target)


def oracle_step(a, p, opt, forward):
    opt.zero_grad()
    out, target = forward()
    loss = loss_of(a, out, target)
    loss.backward()
    opt.step()
    return float(loss.detach())


def cpu_baseline(a, steps=1, warmup=0, n_graphs=None, keep_first=False):
    """Times fwd + loss + bwd + Adamax of the dense port on `n_graphs` graphs of the workload.  keep_first: also
    return (out, {name: grad}) of the very first step, before any parameter update - the parity reference."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = a.cpu_sample if n_graphs is None else n_graphs
    O, p, opt, forward, info = oracle_setup(a, n)
    first = None
    if keep_first:
        opt.zero_grad()
        out, target = forward()
        loss_of(a, out, target).backward()
        first = (out.detach().clone(), {k: v.grad.detach().clone() for k, v in p.items()},
                 {k: v.detach().clone() for k, v in p.items()})
    for _ in range(warmup):
        oracle_step(a, p, opt, forward)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle_step(a, p, opt, forward)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    dense = (" (dense operators: WL is %d MB per graph)" % info["dense_mb"]) if info["dense_mb"] else ""
    res = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d graph(s) of the same workload per step%s, %d timed step(s) of fwd+loss+bwd+Adamax after %d "
                     "warm-up, operators prebuilt" % (n, dense, steps, warmup),
           "ms_per_step": dt * 1e3, "graphs_per_step": n}
    return (res, first) if keep_first else res


def cpu_extrapolated(a):
    """c4: the reference cannot run N = 10 000 (WL is ~120 GB per graph): one graph at N = 250 / 500 /
    1000, power-law fit of seconds per graph, extrapolated to N = a.nodes.  Reported as such."""
    import copy
    import math
    pts = []
    for n in (250, 500, 1000):
        b = copy.copy(a)
        b.nodes = n
        r = cpu_baseline(b, steps=1, warmup=0, n_graphs=1)
        pts.append((n, 1.0 / r["value"]))
    xs = [math.log(n) for n, _ in pts]
    ys = [math.log(t) for _, t in pts]
    mx, my = sum(xs) / 3, sum(ys) / 3
    slope = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sum((x - mx) ** 2 for x in xs)
    t_big = math.exp(my + slope * (math.log(a.nodes) - mx))
    return {"value": 1.0 / t_big, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "reference infeasible at N=%d (dense WL ~120 GB per graph); extrapolated from 1 graph at "
                      "N=250/500/1000 (%.2f / %.2f / %.2f s per step), fitted exponent %.2f"
                      % (a.nodes, pts[0][1], pts[1][1], pts[2][1], slope)}


def run_reference(a):
    """--impl reference: the CPU port alone, K timed steps after W warm-ups, rank 0 only.  The sample per step is the
    largest of (configured batch, 4, 2, 1 graphs) whose K + W steps fit ~150 s, estimated from one timed step of the
    smallest candidate; the configuration printed is the one that ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.config == "c4":
        cb = cpu_extrapolated(a)
        cb["ms_per_step"] = 1e3 / cb["value"]
        n = 1
    else:
        full = graphs_per_rank(a, 1)
        cands = sorted({c for c in (full, 4, 2, 1) if c <= full}, reverse=True)
        probe = cpu_baseline(a, steps=1, warmup=0, n_graphs=cands[-1])
        per_graph = probe["ms_per_step"] / 1e3 / cands[-1]
        budget = 150.0
        n = cands[-1]
        for c in cands:
            if per_graph * c * (a.steps + a.warmup + 1) <= budget:
                n = c
                break
        cb = cpu_baseline(a, steps=a.steps, warmup=a.warmup, n_graphs=n)
        cb["sample"] += "; batch ladder %s tried against a %.0f s budget, %d graph(s) per step ran" % (cands, budget, n)
    line = {"impl": "reference", "metric": a.metric, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a, graphs_per_step=n),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = pure-Python CPU code; timed through the oracle port of its dense torch.mm path on the "
                    "host cores (the checkout cannot travel to the GPU box).  config.graphs_per_step is the sample that "
                    "was timed: batch-norm statistics and per-step overheads are those of that sample, the GPU arm "
                    "runs config.graphs_per_gpu graphs per step"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler(object):
    """SM clocks and throttle reasons DURING the timed region: ONE `nvidia-smi -lms 200` process (the profiling recipe's
    clocks line), started by local rank 0 for every GPU of the job well before the timed region (NVML initialisation on
    an 8-GPU host takes longer than the region itself) and killed after it; only the samples whose time stamps fall
    inside the region count.  (A sampler per rank that spawned nvidia-smi every 100 ms put eight NVML initialisations
    on the host at once, inside the timed region.)"""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, n_gpus=1):
        self.samples, self.proc, self.t0 = [], None, None
        self.ids = ",".join(str(i) for i in range(n_gpus)) if index == 0 else None

    def start(self):
        if self.ids is not None:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", self.ids, "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                import atexit
                atexit.register(self._kill)      # the loop never ends by itself: no orphan if the run dies early
            except Exception:
                self.proc = None
        return self

    def _kill(self):
        if self.proc is not None:
            try:
                self.proc.kill()
            except Exception:
                pass

    def __enter__(self):          # the timed region begins
        import datetime
        self.t0 = datetime.datetime.now()
        return self

    def __exit__(self, *exc):     # ... and ends
        import datetime
        t1 = datetime.datetime.now()
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=6)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
            out = ""
        self.proc = None
        slack = datetime.timedelta(milliseconds=50)
        for ln in (out or "").splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
            except ValueError:
                continue
            if self.t0 - slack <= ts <= t1 + slack:
                self.samples.append(f[1:])

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def gnn_layer_algorithmic_bytes(pack, F):
    """GNN layer: gmul(A) with K output blocks, one pass."""
    Rn, K, nnzA = pack.Rn, pack.K, pack.a[0].nnz
    return 4 * (Rn + 1) + 8 * nnzA + 4 * Rn + 4 * Rn * F + 4 * Rn * K * F


def ccn_level_algorithmic_bytes(st, C, h):
    """CCN level: per vertex i read sum_{j in N(i)} d_j^2 C 4 (gathered tiles) + 4 sum_j d_j (neighbour
    lists), write d_i^2 h 4; weights 18 C h 4 once."""
    d = (st.nbr_ptr[1:] - st.nbr_ptr[:-1]).to(torch.float64)
    dn = d[st.nbr.long()]
    read = float((dn * dn).sum()) * C * 4 + 4.0 * float(dn.sum())
    write = float((d * d).sum()) * h * 4
    return int(read + write + 18 * C * h * 4)


def lgnn_layer_algorithmic_bytes(pack, F):
    """one sparse operator application Y = S X moves 4(R_out+1) [rowptr] + 8 nnz
    [col+val] + 4 R_in F [features once] + 4 R_out K_out F [write]; diagonal operators add 4 R_out;
    the Pm/Pd pair costs 12 B per non-zero.  Returns (edge side, node side) bytes of ONE forward
    application with feature width F on both node and edge states."""
    Rn, Rm, K = pack.Rn, pack.Rm, pack.K
    nnzA, nnzB, nnzP = pack.a[0].nnz, pack.b[0].nnz, pack.p.nnz
    gmul_a = 4 * (Rn + 1) + 8 * nnzA + 4 * Rn + 4 * Rn * F + 4 * Rn * K * F
    gmul_b = 4 * (Rm + 1) + 8 * nnzB + 4 * Rm + 4 * Rm * F + 4 * Rm * K * F
    pmul_n = 4 * (Rn + 1) + 12 * nnzP + 4 * Rm * F + 4 * Rn * 2 * F       # Pm/Pd . XL  -> nodes
    pmul_e = 4 * (Rm + 1) + 12 * nnzP + 4 * Rn * F + 4 * Rm * 2 * F       # Pm^T/Pd^T . X -> edges
    return gmul_b + pmul_e, gmul_a + pmul_n


def profile_step(train_step, resident, flush, reps=20):
    """Device time of every C-ABI entry point of a real training step.

    One eager step is recorded (name, side, ctypes args of every call).  Each distinct (entry point,
    side kind) is then re-issued `reps` times inside a CUDA graph - no host gaps - once back to back
    (warm L2) and once with a 256 MiB L2-flush write before every launch (cold; the flush-only graph
    is timed separately and subtracted).  CUDA events on the launching stream.  Re-issuing a call
    repeats its accumulator atomics: numerically meaningless, identical work.
    Returns {(name, kind): dict(n=launches per step, warm_us=..., cold_us=...)}."""
    from hgnn_b200 import _lib, engine, ops
    calls = []
    orig = _lib.call

    def rec(name, *args):
        calls.append((name, _lib.tag, args))
        return orig(name, *args)

    _lib.call = engine.call = ops.call = rec
    use_program = engine.USE_PROGRAM
    engine.USE_PROGRAM = False      # the per-side Python loop issues the same launches one visible call at a time
    try:
        train_step(resident)
    finally:
        _lib.call = engine.call = ops.call = orig
        engine.USE_PROGRAM = use_program
    torch.cuda.synchronize()

    def kind_of(name, tag):
        if not name.startswith("hgnn_lg_side"):
            return ""
        k = "edge" if tag.endswith(".edge") else "node" if tag.endswith(".node") else tag
        return ("layer0." + k) if tag.startswith("L0.") else k

    cap_stream = torch.cuda.Stream()

    def graph_time(body):
        # manual capture: torch.cuda.graph() would empty the allocator cache on entry and thereby
        # unmap the (already freed, still cached) activation buffers the recorded launches point to
        g = torch.cuda.CUDAGraph()
        cap_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap_stream):
            g.capture_begin()
            for _ in range(reps):
                body()
            g.capture_end()
        torch.cuda.current_stream().wait_stream(cap_stream)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    flush_us = graph_time(lambda: flush.zero_())
    out = {}
    for name, tag, args in calls:
        key = (name, kind_of(name, tag))
        if key in out:
            out[key]["n"] += 1
            continue
        fn = getattr(_lib.lib, name)
        st = torch.cuda.current_stream

        def launch(fn=fn, args=args):
            # the stream argument is the last one: re-target it at the capturing stream
            fn(*(args[:-1] + (st().cuda_stream,)))

        def cold(launch=launch):
            flush.zero_()
            launch()

        out[key] = {"n": 1, "warm_us": graph_time(launch), "cold_us": max(graph_time(cold) - flush_us, 0.0)}
    return out


def _trace(msg):
    if os.environ.get("HGNN_BENCH_TRACE"):
        print("[rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# workloads of our arm: host instances -> device batch -> model output
# --------------------------------------------------------------------------------------------
class GnnWork(object):
    """GNN_simple / GNN_lg on SBM or QM9-shaped graphs through prepare_batch and the reference-shaped models."""

    def __init__(self, a, dev, rank, world, n_host_batches=2):
        from hgnn_b200 import synth
        from hgnn_b200.models.gnns.model_mnb import GNN_lg, GNN_simple
        self.a, self.dev = a, dev
        per = graphs_per_rank(a, world)
        self.per = per

        def make(first_id):
            if a.data == "sbm":
                return synth.sbm_dataset(per, N=a.nodes, a=a.sbm_a, b=a.sbm_b, J=a.J, sparse=True, first_id=first_id)
            return synth.qm9_shaped_dataset(per, J=a.J, sparse=True, first_id=first_id)
        self.host_batches = [make((rank * n_host_batches + k) * per) for k in range(n_host_batches)]
        torch.manual_seed(0)
        if a.kind == "lgnn":
            self.model = GNN_lg(0, a.h, a.layers, 5, a.dim_out, a.J, a.order).to(dev).train()
        else:
            self.model = GNN_simple(0, a.h, a.layers, 5, a.dim_out, a.J).to(dev).train()

    def prepare(self, k):
        from hgnn_b200.functions.batching import prepare_batch
        return prepare_batch(self.host_batches[k % len(self.host_batches)], 0, self.a.J)

    def to_device(self, batch):
        X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = batch
        target = T if self.a.loss == "mse" else T.squeeze(1).long()
        staged = [X.pin_memory(), XL.pin_memory(), target.pin_memory()]
        Xd, XLd, yd = [t.to(self.dev, non_blocking=True) for t in staged]
        h2d = W.pack.nbytes + sum(t.numel() * t.element_size() for t in staged)
        return (Xd, XLd, W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, yd), h2d

    def forward(self, db):
        Xd, XLd, W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, yd = db
        if self.a.kind == "lgnn":
            return self.model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg), yd
        return self.model([Xd, W], N_batch, mask), yd

    def loader_batches(self, n_steps):
        from hgnn_b200.functions.batching import BatchLoader
        data = [inst for hb in self.host_batches for inst in hb]
        nb = len(self.host_batches)
        idx = [list(range((k % nb) * self.per, (k % nb + 1) * self.per)) for k in range(n_steps)]
        return BatchLoader(data, idx, 0, self.a.J, device=self.dev)

    def describe(self, db):
        pack = db[2].pack
        d = {"rows_nodes": pack.Rn, "nnz_A": pack.a[0].nnz, "bytes": pack.nbytes}
        if self.a.kind == "lgnn":
            d.update({"rows_line_graph": pack.Rm, "active_line_graph_rows": int(pack.erow.numel()), "nnz_B": pack.b[0].nnz,
                      "nnz_B_collapsed_T": pack.btc.nnz, "nnz_P": pack.p.nnz})
        return d

    def side_bytes(self, db):
        pack = db[2].pack
        if self.a.kind == "lgnn":
            e, n = lgnn_layer_algorithmic_bytes(pack, 2 * self.a.h)
            return {"edge": e, "node": n}
        return {"node": gnn_layer_algorithmic_bytes(pack, 2 * self.a.h)}

    def load_oracle_params(self, params):
        self.model.load_state_dict({k: v for k, v in params.items()})


class CcnWork(object):
    """CCN_2D on QM9-shaped graphs, a.bs graphs per step in one launch group (CCN_2D.forward_batch)."""

    def __init__(self, a, dev, rank, world, n_host_batches=2):
        from hgnn_b200 import synth
        from hgnn_b200.models.compnets.model_ccn import CCN_2D
        self.a, self.dev, self.per = a, dev, graphs_per_rank(a, world)
        self.host_batches = []
        for k in range(n_host_batches):
            insts = synth.qm9_shaped_dataset(self.per, J=1, sparse=True, first_id=(rank * n_host_batches + k) * self.per)
            self.host_batches.append([(i[0], i[1] + torch.eye(i[1].shape[0]), i[2][:1]) for i in insts])   # train_ccn.py:36
        torch.manual_seed(0)
        self.model = CCN_2D(5, a.dim_out, a.h, a.layers, True).to(dev).train()

    def prepare(self, k):
        hb = self.host_batches[k % len(self.host_batches)]
        X = torch.cat([x for x, _, _ in hb], 0)
        y = torch.stack([t for _, _, t in hb])
        return hb, X, y

    def to_device(self, batch):
        from hgnn_b200.functions.utils_ccn import CcnStructure
        hb, X, y = batch
        st = CcnStructure.from_graphs([A for _, A, _ in hb], device=self.dev)      # neighbour lists: host -> device
        Xp, yp = X.pin_memory(), y.pin_memory()
        Xd, yd = Xp.to(self.dev, non_blocking=True), yp.to(self.dev, non_blocking=True)
        h2d = X.numel() * 4 + y.numel() * 4 + int(st.nbr.numel() + st.nbr_ptr.numel()) * st.nbr.element_size()
        return (Xd, st, yd), h2d

    def forward(self, db):
        Xd, st, yd = db
        rows = st.row_vertex2
        return self.model.fc(self.model._levels(Xd.index_select(0, rows), st)), yd

    def loader_batches(self, n_steps):
        return None

    def describe(self, db):
        st = db[1]
        return {"vertices": st.V, "max_receptive_field": st.nmax, "neighbour_entries": int(st.nbr.numel())}

    def side_bytes(self, db):
        return {"": ccn_level_algorithmic_bytes(db[1], 5, self.a.h)}

    def load_oracle_params(self, params):
        self.model.load_state_dict({k: v for k, v in params.items()})


def parity_check(a, work_cls, dev):
    """The CUDA model against the CPU oracle on the sample the cpu_baseline leg builds: same seeded graphs, the
    oracle's parameters copied in, output and every parameter gradient.  Returns (parity dict, cpu_baseline dict)."""
    import copy
    if a.config == "c4":
        return None, cpu_extrapolated(a)
    n = max(1, min(a.cpu_sample, graphs_per_rank(a, 1)))
    cb, (o_out, o_grads, o_params) = cpu_baseline(a, steps=1, warmup=0, n_graphs=n, keep_first=True)
    b = copy.copy(a)
    b.bs, b.scaling = n, "weak"
    w = work_cls(b, dev, 0, 1, n_host_batches=1)
    w.load_oracle_params(o_params)
    db, _ = w.to_device(w.prepare(0))
    out, target = w.forward(db)
    loss_of(a, out, target).backward()
    torch.cuda.synchronize()

    def rel(x, y, floor=0.0):
        x, y = x.detach().double().cpu(), y.double()
        return float((x - y).abs().max() / max(float(y.abs().max()), floor, 1e-30))
    gmax = max(float(g.abs().max()) for g in o_grads.values())
    e_out = rel(out.view(-1), o_out.view(-1))
    e_grad, worst = 0.0, ""
    for k, v in w.model.named_parameters():
        e = rel(v.grad, o_grads[k], 0.1 * gmax)     # near-zero gradients are compared at 1e-5 x the largest one
        if e > e_grad:
            e_grad, worst = e, k
    par = {"out_rel_err": e_out, "max_grad_rel_err": e_grad, "worst_grad": worst, "n_graphs": n, "L": a.layers,
           "tolerance": TOL, "ok": bool(e_out < TOL and e_grad < TOL),
           "against": "CPU oracle port (oracle/hgnn_oracle.py) on the first %d graph(s) of the workload, same parameters" % n}
    return par, cb


def gpu_dense_baseline(a):
    """The reference's dense code path with .cuda() on this GPU (second baseline line): the
    oracle port - plain torch ops on the zero-padded dense operators - on a small sample."""
    if a.config == "c4" or a.kind == "ccn2":
        return None
    n = max(1, min(a.cpu_sample, graphs_per_rank(a, 1)))
    try:
        O, p, opt, forward, info = oracle_setup(a, n, device="cuda")
        for _ in range(2):
            oracle_step(a, p, opt, forward)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps = 5
        for _ in range(steps):
            oracle_step(a, p, opt, forward)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        return {"value": n / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "graphs_per_step": n,
                "what": "the dense reference path (oracle port: per-graph torch.mm over zero-padded W / WL / Pm / Pd) "
                        "moved to the GPU with .cuda(), %d graph(s) per step, eager, wall clock" % n}
    except Exception as e:       # e.g. out of memory on the dense operators
        return {"value": None, "error": str(e)[:200]}


def run_ours(a):
    import torch.distributed as dist
    import hgnn_b200
    from hgnn_b200.dist import FlatParams, FusedAdamax

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != a.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (a.gpus, world), file=sys.stderr)

    work_cls = CcnWork if a.kind == "ccn2" else GnnWork
    work = work_cls(a, dev, rank, world)
    per = work.per
    model = work.model
    fp = FlatParams(model)
    _trace("model built, broadcasting parameters")
    fp.broadcast(0)
    _trace("broadcast done")
    opt = FusedAdamax(fp, lr=1e-3)
    sampler = ClockSampler(local, world).start()      # NVML is up long before the timed region

    def train_step(dbatch, collective=True):
        fp.zero_grad()
        out, target = work.forward(dbatch)
        loss = loss_of(a, out, target)
        loss.backward()
        if collective:
            fp.all_reduce_grad()        # the only collective of the step (NCCL, or fused into opt.step over peer memory)
        else:
            fp.gather_grad()
        opt.step(grad_scale=1.0 / world, collective=collective)
        return loss

    resident, h2d_bytes = work.to_device(work.prepare(0))
    torch.cuda.synchronize()
    _trace("batch resident, eager warm-up")

    # ---- warm-up (eager), then capture the whole step in a CUDA graph
    launches0 = hgnn_b200.launch_count()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(3, a.warmup)):
            loss = train_step(resident)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    launches_per_step = (hgnn_b200.launch_count() - launches0) // max(3, a.warmup)
    graph, graph_nocoll = None, None
    _trace("eager warm-up done, capturing the step")
    if not a.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = train_step(resident)
        except Exception as e:
            if world > 1:
                raise
            print("warning: CUDA-graph capture failed (%s); timing eager launches" % str(e)[:120], file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    _trace("capture done")

    def step():
        if graph is not None:
            graph.replay()
            return static_loss
        return train_step(resident)

    for _ in range(a.warmup):
        step()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    events = []
    with sampler as clocks:
        for _ in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = step()
            e1.record()
            events.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # keep the GPU busy a little longer so that the clock sampler sees it under load.  A FIXED number
        # of extra steps: every replay contains the gradient all-reduce, so all ranks must issue the
        # same count (a time-based loop would leave unmatched collectives behind).
        for _ in range(700):
            step()
        torch.cuda.synchronize()
    elapsed = sum(e0.elapsed_time(e1) for e0, e1 in events) * 1e-3
    t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
    per_rank_ms = None
    if world > 1:
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)                      # every rank's own device time: shows the skew the MAX hides
        per_rank_ms = [round(float(x.item()) * 1e3 / a.steps, 4) for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    _trace("timed region done")
    value = per * world * a.steps / elapsed
    final_loss = float(loss.item())

    # ---- do all ranks hold the same parameters after the same number of steps?  (flat buffer checksums, all-gathered)
    ranks_agree, exposed_us = None, None
    if world > 1:
        chk = torch.stack([fp.flat.double().sum(), fp.flat.double().abs().sum(), fp.flat.double().pow(2).sum()])
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        ranks_agree = bool(all(torch.equal(c, allc[0]) for c in allc))
        # exposure of the gradient all-reduce: the same captured step without it (parameters diverge from here on;
        # nothing below compares them)
        if graph is not None:
            graph_nocoll = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_nocoll):
                train_step(resident, collective=False)
            for g in (graph, graph_nocoll):
                for _ in range(3):
                    g.replay()
            torch.cuda.synchronize()
            tt = []
            for g in (graph, graph_nocoll):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    g.replay()
                e1.record()
                e1.synchronize()
                tt.append(e0.elapsed_time(e1) * 1e3 / 50)
            d = torch.tensor([tt[0] - tt[1]], device=dev, dtype=torch.float64)
            dist.all_reduce(d, op=dist.ReduceOp.MAX)
            exposed_us = float(d.item())

    # ---- end to end through the public API: host instances -> prepare_batch -> step -> loss.item()
    e2e = None
    if not a.skip_e2e:
        e2e_steps = 50 if a.steps >= 20 else max(3, a.steps)    # ~0.1 s per loop: one host hiccup must not dominate
        if a.config == "c4":
            e2e_steps = min(e2e_steps, 10)
        for k in range(6):      # warm-up: pinned slabs / staging slots / allocator pools reach steady state
            db, _ = work.to_device(work.prepare(k))
            train_step(db).item()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        h2d_tot = 0
        for k in range(e2e_steps):
            db, nb = work.to_device(work.prepare(k))
            h2d_tot += nb
            train_step(db).item()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": per * world * e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d_tot // e2e_steps, "d2h_bytes_per_step": 4,
               "steps": e2e_steps, "ms_per_step": float(dt.item()) * 1e3 / e2e_steps,
               "path": "the reference's loop shape (scripts/train_mnb.py:43-70 / train_ccn.py:31-71), synchronous: host "
                       "instances -> batch preparation -> pinned H2D -> model fwd -> loss -> bwd -> all-reduce -> fused "
                       "Adamax -> loss.item()"}
        loader = work.loader_batches(e2e_steps + 10)
        if loader is not None:
            # same loop fed by functions.batching.BatchLoader (prepare_batch of batch k+1 on a background thread + copy
            # stream while batch k trains); every step still copies its inputs from pinned host memory and reads the loss
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = None
            for k, batch in enumerate(loader):
                if k == 10:     # pinned pools of the producer thread, copy-stream allocator pool: steady state
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                db, _ = work.to_device(batch)
                train_step(db).item()
            torch.cuda.synchronize()
            dt2 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
            e2e["prefetch"] = {"value": per * world * e2e_steps / float(dt2.item()), "unit": UNIT,
                               "ms_per_step": float(dt2.item()) * 1e3 / e2e_steps,
                               "path": "same loop over functions.batching.BatchLoader (one batch of look-ahead)"}

    def finish():
        """End of the run for world > 1.  The captured CUDA graph holds NCCL kernels, and tearing the
        process group down under it can hang, so: rank 0 announces completion through the rendezvous
        store (no NCCL), everybody leaves with os._exit."""
        if world == 1:
            return
        import datetime
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("bench_done", "1")
        else:
            store.wait(["bench_done"], datetime.timedelta(minutes=15))
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        finish()
        return

    # ---- roofline of the dominant aggregation kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    prof = profile_step(lambda b: train_step(b, collective=False), resident, flush)   # rank 0 only: no collective
    side_bytes = work.side_bytes(resident)
    step_us = sum(v["n"] * v["warm_us"] for v in prof.values())
    breakdown = sorted(([k[0] + ("[" + k[1] + "]" if k[1] else ""), v["n"], round(v["warm_us"], 2),
                         round(v["cold_us"], 2), round(100 * v["n"] * v["warm_us"] / step_us, 1)]
                        for k, v in prof.items()), key=lambda r: -r[4])
    # the dominant kernel = the aggregation entry point / side with the largest share of the step
    if a.kind == "ccn2":
        cand = [k for k in prof if k[0].startswith("hgnn_ccn")]
    else:
        cand = [k for k in prof if k[1] in side_bytes]
    dom = max(cand, key=lambda k: prof[k]["n"] * prof[k]["warm_us"])
    dom_bytes = side_bytes[dom[1] if a.kind != "ccn2" else ""]
    t_cold, t_warm = prof[dom]["cold_us"] * 1e-6, prof[dom]["warm_us"] * 1e-6
    achieved = dom_bytes / t_cold / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = a.config if a.config != "c2" else ("h%d" % a.h if a.h != 2 else None)
        traffic = (tj if key is None else tj.get(key, {})).get("%s[%s]" % dom if dom[1] else dom[0])
    except Exception:
        pass
    if a.kind == "ccn2":
        note = ("one CTA per vertex; receptive-field tiles gathered through an index map, 8 partial sums in shared memory "
                "serve all 18 contractions; V = %d vertices of degree <= %d: latency-bound" % (resident[1].V, resident[1].nmax))
        kernel = "%s = CCN-2 promote + 18 contractions + Linear + ReLU per vertex (its backward for _bwd)" % dom[0]
    elif a.h < 16:
        note = ("h=%d: <= %d MB per launch, L2-resident; thread-per-row kernels over the collapsed line graph (one "
                "representative per block of identical phantom rows): the fraction is bounded by latency and instruction "
                "issue, not by HBM bandwidth (profiles/README.md)" % (a.h, dom_bytes // 1000000))
        kernel = ("%s [%s side of a middle layer] = fused multi-operator + Pm/Pd gather, conv, ReLU, BN "
                  "(its transposed-gather backward for _bwd)" % dom)
    else:
        note = ("h=%d: tensor-core tile kernels (csrc/engine_wide.cuh / engine_tc5.cuh): staged CSR structure + 3xTF32 "
                "contractions; profiles/README.md" % a.h)
        kernel = ("%s [%s side of a middle layer] = fused multi-operator + Pm/Pd gather, conv, ReLU, BN "
                  "(its transposed-gather backward for _bwd)" % dom)
    roofline = {"bound": "hbm", "kernel": kernel,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": dom_bytes,
                "launch_us_cold_l2": t_cold * 1e6, "launch_us_warm_l2": t_warm * 1e6,
                "achieved_warm_l2": dom_bytes / t_warm / 1e9, "frac_warm_l2": dom_bytes / t_warm / 1e9 / peak,
                "share_of_step_pct": round(100 * prof[dom]["n"] * prof[dom]["warm_us"] / step_us, 1),
                "how": "the recorded launch re-issued 20x inside a CUDA graph, CUDA events on the launching "
                       "stream; `achieved` uses the cold-L2 time (256 MiB flush write before every launch, "
                       "flush-only graph subtracted), `achieved_warm_l2` the back-to-back time (what the launch sees "
                       "inside a step: its inputs were just written by the previous side); algorithmic bytes per "
                       "SURVEY.md 8(d) on the reference's operators (phantom rows included), backward = forward",
                "per_kernel": {"columns": ["entry point [side]", "launches/step", "warm us", "cold us", "% of step"],
                               "rows": breakdown[:10]},
                "note": note}
    line = {"metric": a.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed * 1e3 / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": launches_per_step * a.steps, "launches_per_step": launches_per_step,
            "cuda_graph": graph is not None, "final_loss": final_loss, "roofline": roofline,
            "pack": work.describe(resident)}
    if per_rank_ms is not None:
        line["ms_per_step_by_rank"] = per_rank_ms
    if ranks_agree is not None:
        line["ranks_agree"] = ranks_agree
        line["allreduce_exposed_us"] = exposed_us
        line["allreduce"] = ("fused into the Adamax launch over NVLink peer memory (csrc/p2p.cu)" if opt.peers is not None
                             else "NCCL all-reduce of the flat gradient")
        if opt.peers is not None:
            line["peer_fault"] = int(opt.fault.item())
    failed = False
    if not a.skip_cpu:
        parity, cb = parity_check(a, work_cls, dev)
        line["cpu_baseline"] = {k: v for k, v in cb.items() if k not in ("ms_per_step", "graphs_per_step")}
        if parity is not None:
            line["parity"] = parity
            failed = not parity["ok"]
        dense = gpu_dense_baseline(a)
        if dense is not None:
            line["gpu_dense_baseline"] = dense
    print(json.dumps(line), flush=True)
    if failed:
        print("bench.py: PARITY FAILURE against the CPU oracle: %s" % json.dumps(line["parity"]), file=sys.stderr, flush=True)
        if world == 1:
            sys.exit(1)
    finish()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(a)


if __name__ == "__main__":
    main()
