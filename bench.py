#!/usr/bin/env python
"""bench.py -- LGNN training throughput on synthetic binary-SBM graphs (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one LGNN training step (forward + cross-entropy + backward + gradient all-reduce +
Adamax) over one batch of 32 SBM graphs with N=1000 nodes per GPU (weak scaling: every rank owns its
own 32 graphs; the only collective is the flat-gradient all-reduce).  Model =
``GNN_lg(task=0, h=2, L=20, dim_input=5, dim_output=2, J=1, order=1)`` - the reference's script
defaults (scripts/main_gnn.py:59,75-77), fp32.

Printed JSON (one line, rank 0):
  value   graphs/s with the batch already packed in HBM (CUDA-graph replay of the whole step, CUDA
          events per step, L2 flushed between steps, max over ranks);
  e2e     graphs/s through the public API from HOST instances: prepare_batch (per-graph DMA out of the
          pinned dataset + GPU gather into block-diagonal CSR) -> model -> loss -> backward -> optimizer
          -> loss.item(); e2e.prefetch = the same loop fed by functions.batching.BatchLoader;
  roofline  the dominant aggregation kernel (fused edge-side update) timed alone with CUDA events,
          algorithmic bytes per SURVEY.md 8(d) / DESIGN.md, against MEASURED_PEAKS.json;
  cpu_baseline  the CPU oracle port (oracle/hgnn_oracle.py = the reference's dense torch.mm loops)
          timed on the host cores on a bounded sample of the same workload.
``--impl reference`` times that CPU port alone (the reference is pure Python and cannot travel to
the GPU box).
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

METRIC = "lgnn_sbm_train_graphs_per_s"
UNIT = "graphs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bs", type=int, default=32, help="graphs per GPU")
    ap.add_argument("--nodes", type=int, default=1000)
    ap.add_argument("--h", type=int, default=2)
    ap.add_argument("--layers", type=int, default=20)
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--J", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=2, help="graphs per CPU-baseline step")
    return ap.parse_args()


def workload_config(a):
    return {"workload": "LGNN (GNN_lg order %d, L=%d, h=%d, J=%d) on 2-class binary SBM N=%d (a=7,b=3), "
                        "batch %d graphs per GPU, fp32 fwd+bwd+Adamax" % (a.order, a.layers, a.h, a.J, a.nodes, a.bs),
            "graphs_per_gpu": a.bs, "nodes_per_graph": a.nodes, "layers": a.layers, "h": a.h, "J": a.J,
            "order": a.order, "parallelism": "dp%d (graphs sharded, flat-gradient all-reduce)" % a.gpus,
            "l2": "256 MiB buffer written between timed steps (L2 flush)"}


# --------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's dense path, on the host cores
# --------------------------------------------------------------------------------------------
def cpu_reference_setup(a, n_graphs, first_id=0):
    from oracle import hgnn_oracle as O
    from hgnn_b200 import synth
    inst = []
    for i in range(n_graphs):
        s = synth.sbm_instance(first_id + i, N=a.nodes, J=a.J, sparse=True)
        A = s[1].to_dense()
        inst.append([s[0], A, s[2]] + list(O.graph_operators([s[0], A], a.J, True)))
    batch = O.prepare_batch(inst, 0, a.J)
    p = O.init_gnn_params("lg", a.h, a.layers, 5, 2, a.J, a.order, seed=0)
    for v in p.values():
        v.requires_grad_()
    labels = torch.tensor([int(i[2][0]) for i in inst])
    opt = torch.optim.Adamax(list(p.values()), lr=1e-3)
    return O, batch, p, labels, opt


def cpu_reference_step(a, O, batch, p, labels, opt):
    X, W, _, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = batch
    opt.zero_grad()
    out = O.gnn_lg_forward(p, a.layers, a.order, [X, XL, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
    loss = torch.nn.functional.cross_entropy(out, labels)
    loss.backward()
    opt.step()
    return float(loss.detach())


def cpu_baseline(a, steps=1, warmup=0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = cpu_reference_setup(a, a.cpu_sample)
    for _ in range(warmup):
        cpu_reference_step(a, *state)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(a, *state)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return {"value": a.cpu_sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d graph(s) of the same SBM N=%d workload per step (dense operators: WL is "
                      "%d MB per graph), %d timed step(s) of fwd+loss+bwd+Adamax, operators prebuilt"
                      % (a.cpu_sample, a.nodes, int(state[1][4][0].numel() * 4 / 1e6), steps),
            "ms_per_step": dt * 1e3}


def run_reference(a):
    """--impl reference: the CPU port alone, K timed steps after W warm-ups, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(a, steps=a.steps, warmup=a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = pure-Python CPU code; timed through the oracle port of its dense "
                    "torch.mm path on the host cores (the checkout cannot travel to the GPU box)"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
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
    from hgnn_b200 import _lib, engine
    calls = []
    orig = _lib.call

    def rec(name, *args):
        calls.append((name, _lib.tag, args))
        return orig(name, *args)

    _lib.call = engine.call = rec
    use_program = engine.USE_PROGRAM
    engine.USE_PROGRAM = False      # the per-side Python loop issues the same launches one visible call at a time
    try:
        train_step(resident)
    finally:
        _lib.call = engine.call = orig
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


def run_ours(a):
    import torch.distributed as dist
    import hgnn_b200
    from hgnn_b200 import _lib, synth
    from hgnn_b200.dist import FlatParams, FusedAdamax
    from hgnn_b200.functions.batching import prepare_batch
    from hgnn_b200.models.gnns.model_mnb import GNN_lg

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

    # ---- data: every rank owns its own a.bs graphs (weak scaling), two distinct host batches
    n_host_batches = 2
    host_batches = [synth.sbm_dataset(a.bs, N=a.nodes, J=a.J, sparse=True,
                                      first_id=(rank * n_host_batches + k) * a.bs) for k in range(n_host_batches)]
    torch.manual_seed(0)
    model = GNN_lg(0, a.h, a.layers, 5, 2, a.J, a.order).to(dev).train()
    fp = FlatParams(model)
    _trace("model built, broadcasting parameters")
    fp.broadcast(0)
    _trace("broadcast done")
    opt = FusedAdamax(fp, lr=1e-3)

    def to_device(batch):
        X, W, T, XL, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch = batch
        labels = T.squeeze(1).long()
        staged = [X.pin_memory(), XL.pin_memory(), labels.pin_memory()]
        Xd, XLd, yd = [t.to(dev, non_blocking=True) for t in staged]
        h2d = W.pack.nbytes + sum(t.numel() * t.element_size() for t in staged)
        return (Xd, XLd, W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, yd), h2d

    def train_step(dbatch, collective=True):
        Xd, XLd, W, WL, Pm, Pd, mask, mask_lg, N_batch, E_batch, yd = dbatch
        fp.zero_grad()
        out = model([Xd, XLd, W, WL, Pm, Pd], N_batch, mask, E_batch, mask_lg)
        loss = torch.nn.functional.cross_entropy(out, yd)
        loss.backward()
        if collective:
            fp.all_reduce_grad()        # the only collective of the step
        else:
            fp.gather_grad()
        opt.step(grad_scale=1.0 / world)
        return loss

    resident, h2d_bytes = to_device(prepare_batch(host_batches[0], 0, a.J))
    pack = resident[2].pack
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
    graph = None
    _trace("eager warm-up done, capturing the step")
    if not a.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = train_step(resident)
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
    with ClockSampler(local) as clocks:
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
        for _ in range(300):
            step()
        torch.cuda.synchronize()
    elapsed = sum(e0.elapsed_time(e1) for e0, e1 in events) * 1e-3
    t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    _trace("timed region done")
    value = a.bs * world * a.steps / elapsed
    final_loss = float(loss.item())

    # ---- end to end through the public API: host instances -> prepare_batch -> step -> loss.item()
    e2e = None
    if not a.skip_e2e:
        e2e_steps = 50 if a.steps >= 20 else max(3, a.steps)    # ~0.1 s per loop: one host hiccup must not dominate
        for k in range(6):      # warm-up: pinned slabs / staging slots / allocator pools reach steady state
            db, _ = to_device(prepare_batch(host_batches[k % n_host_batches], 0, a.J))
            train_step(db).item()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        h2d_tot = 0
        for k in range(e2e_steps):
            db, nb = to_device(prepare_batch(host_batches[k % n_host_batches], 0, a.J))
            h2d_tot += nb
            train_step(db).item()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": a.bs * world * e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d_tot // e2e_steps, "d2h_bytes_per_step": 4,
               "steps": e2e_steps, "ms_per_step": float(dt.item()) * 1e3 / e2e_steps,
               "path": "the reference's loop shape (scripts/train_mnb.py:43-70), synchronous: prepare_batch(host "
                       "instances) -> pinned H2D -> GNN_lg fwd -> CE loss -> bwd -> all-reduce -> fused Adamax -> "
                       "loss.item()"}
        # same loop fed by functions.batching.BatchLoader (prepare_batch of batch k+1 on a background
        # thread + copy stream while batch k trains); every step still copies its inputs from pinned
        # host memory and reads the loss back
        from hgnn_b200.functions.batching import BatchLoader
        data = [inst for hb in host_batches for inst in hb]
        idx_lists = [list(range((k % n_host_batches) * a.bs, (k % n_host_batches + 1) * a.bs))
                     for k in range(e2e_steps + 10)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = None
        for k, batch in enumerate(BatchLoader(data, idx_lists, 0, a.J, device=dev)):
            if k == 10:     # pinned pools of the producer thread, copy-stream allocator pool: steady state
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            db, _ = to_device(batch)
            train_step(db).item()
        torch.cuda.synchronize()
        dt2 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
        e2e["prefetch"] = {"value": a.bs * world * e2e_steps / float(dt2.item()), "unit": UNIT,
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
    edge_bytes, node_bytes = lgnn_layer_algorithmic_bytes(pack, 2 * a.h)
    step_us = sum(v["n"] * v["warm_us"] for v in prof.values())
    breakdown = sorted(([k[0] + ("[" + k[1] + "]" if k[1] else ""), v["n"], round(v["warm_us"], 2),
                         round(v["cold_us"], 2), round(100 * v["n"] * v["warm_us"] / step_us, 1)]
                        for k, v in prof.items()), key=lambda r: -r[4])
    # the dominant kernel = the entry point / side with the largest share of the step
    dom = max((k for k in prof if k[1] in ("edge", "node")), key=lambda k: prof[k]["n"] * prof[k]["warm_us"])
    dom_bytes = edge_bytes if dom[1] == "edge" else node_bytes
    t_cold, t_warm = prof[dom]["cold_us"] * 1e-6, prof[dom]["warm_us"] * 1e-6
    achieved = dom_bytes / t_cold / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # the captures are per state width (h = 2: thread-per-row kernels, h = 32: tensor-core tile kernels)
        traffic = (tj if a.h == 2 else tj.get("h%d" % a.h, {})).get("%s[%s]" % dom)
    except Exception:
        pass
    if a.h < 16:
        note = ("h=%d: <= %d MB per launch; each launch is a chain of ~5 dependent memory rounds, so the fraction "
                "is bounded by latency, not by HBM bandwidth" % (a.h, dom_bytes // 1000000))
    else:
        note = ("h=%d: tensor-core tile kernels (csrc/engine_wide.cuh): staged CSR structure + 3xTF32 mma.sync "
                "contractions; instruction-issue bound (the on-the-fly fp32 -> 2 x tf32 operand splits), "
                "profiles/README.md" % a.h)
    roofline = {"bound": "hbm",
                "kernel": "%s [%s side of a middle layer] = fused multi-operator + Pm/Pd gather, conv, ReLU, BN "
                          "(its transposed-gather backward for _bwd)" % dom,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": dom_bytes,
                "launch_us_cold_l2": t_cold * 1e6, "launch_us_warm_l2": t_warm * 1e6,
                "achieved_warm_l2": dom_bytes / t_warm / 1e9,
                "share_of_step_pct": round(100 * prof[dom]["n"] * prof[dom]["warm_us"] / step_us, 1),
                "how": "the recorded launch re-issued 20x inside a CUDA graph, CUDA events on the launching "
                       "stream; `achieved` uses the cold-L2 time (256 MiB flush write before every launch, "
                       "flush-only graph subtracted); algorithmic bytes per SURVEY.md 8(d), backward = forward",
                "per_kernel": {"columns": ["entry point [side]", "launches/step", "warm us", "cold us", "% of step"],
                               "rows": breakdown[:10]},
                "note": note}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": elapsed * 1e3 / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a), "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": launches_per_step * a.steps, "launches_per_step": launches_per_step,
            "cuda_graph": graph is not None, "final_loss": final_loss, "roofline": roofline,
            "pack": {"rows_nodes": pack.Rn, "rows_line_graph": pack.Rm, "nnz_A": pack.a[0].nnz,
                     "nnz_B": pack.b[0].nnz, "nnz_P": pack.p.nnz, "bytes": pack.nbytes}}
    if not a.skip_cpu:
        line["cpu_baseline"] = {k: v for k, v in cpu_baseline(a).items() if k != "ms_per_step"}
    print(json.dumps(line), flush=True)
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
