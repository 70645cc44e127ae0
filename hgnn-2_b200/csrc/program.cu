// program.cu -- the whole GNN_simple / GNN_lg layer stack as two host calls.
//
// The Python engine (hgnn-2_b200/engine.py) used to walk the sides of a model itself: ~90 ctypes
// calls per training step, each with a freshly filled descriptor struct - 3 ms of host time against
// 1.3 ms of device time at the headline configuration.  Here the walk is native: the model is
// described once as a static `hgnn_program_t` (tensor table + side list, parameters by index), a
// batch as `hgnn_batch_t`, and hgnn_program_fwd / hgnn_program_bwd issue the very same launches
// (hgnn_lg_side_fwd / hgnn_lg_side_bwd / readout / step-end reductions) from C++.
//
// Reference semantics: models/gnns/model_mnb.py:58-66,124-129 (layer loop), layers_mnb.py:92,:386
// (readout sum over the padded slots).  Host code only; no kernels live in this file.
#include <atomic>
#include <functional>
#include <map>
#include <vector>
#include <cstdlib>
#include "common.cuh"
#include "mega.cuh"

static std::atomic<long long> g_program_launches{0};

extern "C" long long hgnn_program_launches(void) { return g_program_launches.load(); }

namespace {

// activation / gradient workspace: every non-input tensor, then the readout rows (Rn x Fout_readout)
struct WorkLayout {
    std::vector<long long> off;   // per tensor, floats; -1 for the inputs
    long long readout_off = 0;
    long long total = 0;
};

inline long long align32(long long n) { return (n + 31) & ~31ll; }

bool plan_work(const hgnn_program_t* prog, int Rn, int Rm, WorkLayout* w) {
    if (!prog || prog->n_tensors < 1 || prog->n_sides < 1 || !prog->tensors || !prog->sides) return false;
    const int n_in = prog->dual ? 2 : 1;
    w->off.assign(prog->n_tensors, -1);
    long long cur = 0;
    for (int t = n_in; t < prog->n_tensors; ++t) {
        const hgnn_prog_tensor_t& T = prog->tensors[t];
        w->off[t] = cur;
        cur += align32((long long)(T.rows ? Rm : Rn) * T.F);
    }
    const hgnn_prog_side_t& last = prog->sides[prog->n_sides - 1];
    w->readout_off = cur;
    cur += align32((long long)Rn * (last.Ha + last.Hb));
    w->total = cur;
    return true;
}

inline const float* param(const long long* addr, int idx) {
    return idx < 0 ? nullptr : reinterpret_cast<const float*>(static_cast<uintptr_t>(addr[idx]));
}

inline int rows_of(const hgnn_program_t* prog, const hgnn_batch_t* b, int t) {
    return prog->tensors[t].rows ? b->Rm : b->Rn;
}

// how a consumer normalises tensor t on load (training: from the producer's binned sums)
hgnn_bn_ref_t bn_ref(const hgnn_program_t* prog, const hgnn_batch_t* b, int t, const long long* addr,
                     const double* arena) {
    hgnn_bn_ref_t r;
    const hgnn_prog_tensor_t& T = prog->tensors[t];
    r.affine = nullptr;
    if (T.bn_weight < 0) {
        r.acc = nullptr;
        r.weight = r.bias = nullptr;
        r.n_rows = 0;
    } else {
        r.acc = arena + T.acc_f;
        r.weight = param(addr, T.bn_weight);
        r.bias = param(addr, T.bn_bias);
        r.n_rows = rows_of(prog, b, t);
    }
    return r;
}

inline const float* tensor_ptr(const hgnn_program_t* prog, const WorkLayout& w, int t, const float* X,
                               const float* XL, const float* work) {
    if (t == 0) return X;
    if (prog->dual && t == 1) return XL;
    return work + w.off[t];
}

bool check_program(const hgnn_program_t* prog, const hgnn_batch_t* b) {
    for (int i = 0; i < prog->n_sides; ++i) {
        const hgnn_prog_side_t& s = prog->sides[i];
        if (s.src_self < 0 || s.src_self >= prog->n_tensors || s.src_cross >= prog->n_tensors ||
            s.out >= prog->n_tensors || s.Wa < 0 || s.ba < 0 || s.Ha < 1 || (s.Hb > 0 && (s.Wb < 0 || s.bb < 0)))
            return false;
        if ((s.out < 0) != (i == prog->n_sides - 1)) return false;      // exactly the last side is the readout
        if (s.kind == 1 && !prog->dual) return false;
        if (s.src_cross >= 0 && !prog->dual) return false;
    }
    if (prog->dual && (!b->edge_ops || !b->edge_ops_T || !b->p_rowptr || !b->pt_rowptr)) return false;
    return b->node_ops && b->node_ops_T && b->node_off && b->pad_n;
}


// ---- persistent ("mega") kernels: which sides they take, and their parameter block (mega.cuh) ----------
// Opt-in (HGNN_B200_MEGA=1): measured on the C2 workload the persistent kernels are SLOWER than the per-side
// launches (1.39 vs 0.89 ms per step, profiles/README.md "persistent kernels"): a grid barrier plus the reload of
// the batch-norm sums costs ~4 us per side against ~1 us for a programmatic dependent launch, and 16 warps per SM
// running thread-per-row code are bound by instruction latency, not by the launch boundary.
bool mega_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_MEGA"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

bool plain_ops(const hgnn_op_t* ops, int n) {      // [I, D, CSR] without a run-length part
    return ops && n == 3 && ops[0].kind == HGNN_OP_IDENT && ops[1].kind == HGNN_OP_DIAG && ops[1].diag &&
           ops[2].kind == HGNN_OP_CSR && ops[2].rowptr;
}

// sides [s0, s1) run inside the persistent kernels; s0 == s1: none.  Eligible: width-4 inputs and output with
// batch-norm (every middle-layer side at h = 2), J = 1, and for line-graph models the collapsed structure of a
// batch whose edge feature is the line-graph degree (hgnn_batch_t.collapse_ok).
void mega_range(const hgnn_program_t* prog, const hgnn_batch_t* b, int* s0, int* s1) {
    *s0 = *s1 = 0;
    if (mega_disabled() || !b->mega_scratch || b->n_ops != 3 || prog->n_tensors > mk::MAX_TENSORS) return;
    if (!plain_ops(b->node_ops, b->n_ops) || !plain_ops(b->node_ops_T, b->n_ops)) return;
    if (prog->dual) {
        if (!b->collapse_ok || !b->erow || !b->ew || !b->btc_rowptr || !b->btc_col || !b->btc_val) return;
        if (!plain_ops(b->edge_ops, b->n_ops)) return;
    }
    auto ok = [&](int i) {
        const hgnn_prog_side_t& sd = prog->sides[i];
        if (sd.out < 0 || sd.Ha + sd.Hb != 4 || sd.Hb < 1) return false;
        if (prog->tensors[sd.src_self].F != 4 || prog->tensors[sd.out].bn_weight < 0) return false;
        if (sd.src_cross >= 0 && prog->tensors[sd.src_cross].F != 4) return false;
        return true;
    };
    int best0 = 0, best1 = 0;
    for (int i = 0; i < prog->n_sides;) {
        if (!ok(i)) { ++i; continue; }
        int j = i;
        while (j < prog->n_sides && ok(j) && j - i < mk::MAX_SIDES) ++j;
        if (j - i > best1 - best0) { best0 = i; best1 = j; }
        i = j;
    }
    if (best1 - best0 < 2) return;
    *s0 = best0;
    *s1 = best1;
}

void mega_graph(const hgnn_batch_t* b, bool dual, mk::Graph* g) {
    memset(g, 0, sizeof(*g));
    g->Rn = b->Rn;
    g->deg = b->node_ops[1].diag;
    g->a_rp = b->node_ops[2].rowptr; g->a_col = b->node_ops[2].col; g->a_val = b->node_ops[2].val;
    g->at_rp = b->node_ops_T[2].rowptr; g->at_col = b->node_ops_T[2].col; g->at_val = b->node_ops_T[2].val;
    if (!dual) return;
    g->Rm = b->Rm;
    g->n_act = b->n_act;
    g->dl = b->edge_ops[1].diag;
    g->b_rp = b->edge_ops[2].rowptr; g->b_col = b->edge_ops[2].col; g->b_val = b->edge_ops[2].val;
    g->btc_rp = b->btc_rowptr; g->btc_col = b->btc_col; g->btc_val = b->btc_val;
    g->p_rp = b->p_rowptr; g->p_col = b->p_col; g->p_pm = b->p_pm; g->p_pd = b->p_pd;
    g->pt_rp = b->pt_rowptr; g->pt_col = b->pt_col; g->pt_pm = b->pt_pm; g->pt_pd = b->pt_pd;
    g->erow = b->erow;
    g->ew = b->ew;
}

// tensor table + sides [s0, s1) of the program as a kernel parameter block
void mega_params(const hgnn_program_t* prog, const hgnn_batch_t* b, const WorkLayout& w, const float* X, const float* XL,
                 const long long* addr, const float* work, float* gwork, float* gX, double* arena, int s0, int s1,
                 mk::Params* P) {
    mega_graph(b, prog->dual != 0, &P->g);
    P->n_tensors = prog->n_tensors;
    P->n_sides = s1 - s0;
    P->expand = -1;
    P->pad = 0;
    P->bar = static_cast<unsigned int*>(b->mega_scratch);
    P->trace = hgnn_mega_trace_ptr();
    P->nnz1_n = b->node_ops[2].nnz > 0 ? b->node_ops[2].nnz : -1;
    P->nnz1_e = prog->dual && b->edge_ops[2].nnz > 0 ? b->edge_ops[2].nnz : 0;
    P->nnz2 = prog->dual ? b->p_nnz : 0;
    if (prog->dual && (b->edge_ops[2].nnz <= 0 || b->p_nnz <= 0)) P->nnz1_n = -1;
    hgnn_mega_plan(P);
    for (int t = 0; t < prog->n_tensors; ++t) {
        const hgnn_prog_tensor_t& T = prog->tensors[t];
        mk::Tensor& o = P->t[t];
        o.data = const_cast<float*>(tensor_ptr(prog, w, t, X, XL, work));
        o.grad = !gwork ? nullptr : (t == 0 ? gX : (w.off[t] >= 0 ? gwork + w.off[t] : nullptr));
        const bool bn = T.bn_weight >= 0;
        o.acc_f = bn ? arena + T.acc_f : nullptr;
        o.acc_b = bn ? arena + T.acc_b : nullptr;
        o.bn_w = bn ? param(addr, T.bn_weight) : nullptr;
        o.bn_b = bn ? param(addr, T.bn_bias) : nullptr;
        o.n_rows = T.rows ? b->Rm : b->Rn;
        o.pad = 0;
        o.inv_n = o.n_rows > 0 ? 1.0 / (double)o.n_rows : 0.0;
    }
    for (int i = s0; i < s1; ++i) {
        const hgnn_prog_side_t& sd = prog->sides[i];
        mk::Side& o = P->s[i - s0];
        o.kind = sd.kind; o.src_self = sd.src_self; o.src_cross = sd.src_cross; o.out = sd.out;
        o.Wa = param(addr, sd.Wa); o.ba = param(addr, sd.ba); o.Wb = param(addr, sd.Wb); o.bb = param(addr, sd.bb);
        o.Ha = sd.Ha; o.Hb = sd.Hb; o.relu_from = sd.relu_from;
        o.Cin = b->n_ops * 4 + (sd.src_cross >= 0 ? 8 : 0);
        o.dW_bins = arena + sd.dW_off;
        o.db_bins = arena + sd.db_off;
        o.need_self = o.need_cross = o.acc_self = o.acc_cross = 0;
    }
}

// line-graph tensor produced inside [s0, s1) that a per-side kernel outside the range reads on ALL rows:
// -1 none, -2 more than one (the persistent kernels expand a single tensor)
int mega_expand_fwd(const hgnn_program_t* prog, int s0, int s1) {
    int found = -1;
    for (int i = s1; i < prog->n_sides; ++i) {
        const int srcs[2] = {prog->sides[i].src_self, prog->sides[i].src_cross};
        for (int t : srcs) {
            if (t < 0 || !prog->tensors[t].rows) continue;
            bool inside = false;
            for (int k = s0; k < s1; ++k) inside = inside || prog->sides[k].out == t;
            if (!inside || t == found) continue;
            if (found >= 0) return -2;
            found = t;
        }
    }
    return found;
}
// line-graph tensor whose gradient is completed inside the range and consumed by the backward of a side before it
int mega_expand_bwd(const hgnn_program_t* prog, int s0, int s1) {
    int found = -1;
    for (int i = 0; i < s0; ++i) {
        const int t = prog->sides[i].out;
        if (t < 0 || !prog->tensors[t].rows) continue;
        bool inside = false;
        for (int k = s0; k < s1; ++k) inside = inside || prog->sides[k].src_self == t || prog->sides[k].src_cross == t;
        if (!inside || t == found) continue;
        if (found >= 0) return -2;
        found = t;
    }
    return found;
}


// ---- collapsed line graph on the per-side kernels --------------------------------------------------------
// With hgnn_batch_t.collapse_ok the line-graph sides skip the copies of a phantom block (row weight <= 0) and
// weight the representative by its multiplicity; the transposed operator is then the plain CSR btc instead of
// the run-length split.  Only the thread-per-row kernels know row weights, so every side must have one of the
// width combinations they are instantiated for (the feature maps of h = 2: models/gnns/model_mnb.py:48-50,98-100).
bool collapse_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_NO_COLLAPSE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

bool use_collapse(const hgnn_program_t* prog, const hgnn_batch_t* b) {
    if (collapse_disabled() || !prog->dual || !b->collapse_ok || !b->ew || !b->erow || !b->btc_rowptr || !b->btc_col || !b->btc_val) return false;
    if (b->n_ops != 3 || !plain_ops(b->node_ops, 3) || !plain_ops(b->edge_ops, 3) || !plain_ops(b->node_ops_T, 3)) return false;
    const char* e = getenv("HGNN_B200_NO_ROW4");
    if (e && e[0] == '1') return false;
    for (int i = 0; i < prog->n_sides; ++i) {
        const hgnn_prog_side_t& sd = prog->sides[i];
        const int Fs = prog->tensors[sd.src_self].F, Fc = sd.src_cross >= 0 ? prog->tensors[sd.src_cross].F : 0;
        const int Fo = sd.Ha + sd.Hb;
        const bool row4 = Fs == 4 && Fo == 4 && (Fc == 0 || Fc == 4) && sd.out >= 0;
        const bool rowg = (Fs == 5 && Fc == 1 && Fo == 4) || (Fs == 1 && Fc == 4 && Fo == 4) ||
                          (Fs == 4 && Fc == 4 && (Fo == 2 || Fo == 1)) || (Fs == 5 && Fc == 0 && Fo == 4) ||
                          (Fs == 4 && Fc == 0 && (Fo == 2 || Fo == 1));
        if (!row4 && !rowg) return false;
        if (sd.out < 0 && Fo == 4) return false;
    }
    return true;
}

// The one decision both passes share (a forward on the persistent kernels leaves the skipped line-graph rows of
// its activations unwritten, so the backward must take the same path): range, the tensors to expand, and every
// output inside the range consumed by a later side (otherwise its gradient would need a zero fill).
void mega_decide(const hgnn_program_t* prog, const hgnn_batch_t* b, int* m0, int* m1, int* fwd_expand, int* bwd_expand) {
    mega_range(prog, b, m0, m1);
    *fwd_expand = *bwd_expand = -1;
    if (*m1 <= *m0) return;
    *fwd_expand = mega_expand_fwd(prog, *m0, *m1);
    *bwd_expand = mega_expand_bwd(prog, *m0, *m1);
    bool ok = *fwd_expand != -2 && *bwd_expand != -2;
    for (int i = *m0; ok && i < *m1; ++i) {
        bool used = false;
        for (int k = i + 1; k < prog->n_sides; ++k)
            used = used || prog->sides[k].src_self == prog->sides[i].out || prog->sides[k].src_cross == prog->sides[i].out;
        ok = used;
    }
    if (!ok) *m0 = *m1 = 0;
}

}  // namespace

extern "C" int hgnn_program_uses_collapse(const hgnn_program_t* prog, const hgnn_batch_t* b) {
    return (prog && b && prog->sides && prog->tensors && use_collapse(prog, b)) ? 1 : 0;
}

extern "C" long long hgnn_program_work_floats(const hgnn_program_t* prog, int Rn, int Rm) {
    WorkLayout w;
    if (!plan_work(prog, Rn, Rm, &w)) return -1;
    return w.total;
}

// ---------------------------------------------------------------------------------------------------------
// Replay of a pass as ONE graph launch.
//
// A pass issues ~40 dependent side kernels whose sizes and pointers differ from batch to batch, so it cannot be
// captured once and replayed as it is; issuing them costs the host 0.19 ms per pass, and the synchronous training
// loop is host-bound (profiles/logs/e2e_phases_r2k.log).  But the SEQUENCE of kernels is the same for every batch of a
// model.  So the side loop runs with a launch recorder installed (engine.cu: eng_launch appends kernel, launch
// dimensions and the by-value argument block instead of launching), the recorded slots are written into the kernel
// nodes of a graph that was captured once from such a recording (cudaGraphExecKernelNodeSetParams: 0.4 us per node),
// and the graph is launched.  Measured (profiles/graph_update_probe.cu, 40 nodes): 15 us of host time against 75 us
// for 40 launches.  The few launches around the sides (arena memset, readout sum, running statistics, gradient
// reduction) stay direct: before the graph, or deferred until after it.
// ---------------------------------------------------------------------------------------------------------
struct ReplayCtx {
    hgnn_eng_recorder_t rec;
    std::vector<std::function<int()>> post;      // direct launches that must follow the recorded kernels
    bool restartable = true;                     // false once a non-idempotent direct launch was issued
};

struct ReplayGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::vector<cudaGraphNode_t> nodes;
    std::vector<const void*> funcs;
    std::vector<int> block, pdl;
    std::vector<unsigned> smem;
    int failures = 0;
};

// Opt-in (HGNN_B200_REPLAY=1).  Measured on C2 (same box, profiles/logs/e2e_replay_ab_r2w.log): issuing a pass drops
// from 0.35 to 0.27 ms (forward) / 0.41 to 0.36 ms (backward), but the step does not get shorter (1.86 vs 1.86, 2.07 vs
// 1.77 ms): the GPU cannot start a pass before all of its nodes are parametrised and the graph is launched, whereas
// direct launches stream - the device works on side 1 while the host issues side 2 - and at ~5 us per launch against
// 7-13 us per kernel the host stays ahead of the device anyway.  The synchronous loop is a latency chain (prepare ->
// forward on the device -> loss -> backward on the device -> read-back), not an issue-rate problem.
bool replay_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_REPLAY"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

// every side on the thread-per-row kernels (the only ones the recorder sees): the feature maps of h = 2
bool rowpath_program(const hgnn_program_t* prog, const hgnn_batch_t* b) {
    if (b->n_ops != 3 || !plain_ops(b->node_ops, 3) || !b->node_ops_T) return false;
    if (prog->dual && (!plain_ops(b->edge_ops, 3) || !b->edge_ops_T)) return false;
    const char* e = getenv("HGNN_B200_NO_ROW4");
    if (e && e[0] == '1') return false;
    for (int i = 0; i < prog->n_sides; ++i) {
        const hgnn_prog_side_t& sd = prog->sides[i];
        const int Fs = prog->tensors[sd.src_self].F, Fc = sd.src_cross >= 0 ? prog->tensors[sd.src_cross].F : 0;
        const int Fo = sd.Ha + sd.Hb;
        const bool row4 = Fs == 4 && Fo == 4 && (Fc == 0 || Fc == 4) && sd.out >= 0;
        const bool rowg = (Fs == 5 && Fc == 1 && Fo == 4) || (Fs == 1 && Fc == 4 && Fo == 4) ||
                          (Fs == 4 && Fc == 4 && (Fo == 2 || Fo == 1)) || (Fs == 5 && Fc == 0 && Fo == 4) ||
                          (Fs == 4 && Fc == 0 && (Fo == 2 || Fo == 1));
        if (!row4 && !rowg) return false;
        if (sd.out < 0 && Fo == 4) return false;
    }
    return true;
}

bool replay_wanted(const hgnn_program_t* prog, const hgnn_batch_t* b, cudaStream_t s) {
    if (replay_disabled() || !rowpath_program(prog, b)) return false;
    int m0 = 0, m1 = 0, e0 = -1, e1 = -1;
    mega_decide(prog, b, &m0, &m1, &e0, &e1);
    if (m1 > m0) return false;                       // the persistent kernels are their own launch
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return false;                                // inside somebody's capture: issue the launches into it
    }
    return true;
}

void replay_destroy(ReplayGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g.exec = nullptr; g.graph = nullptr;
    g.nodes.clear(); g.funcs.clear(); g.block.clear(); g.pdl.clear(); g.smem.clear();
}

int replay_launch_slot(const hgnn_eng_slot_t& sl, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sl.grid);
    cfg.blockDim = dim3(sl.block);
    cfg.dynamicSmemBytes = sl.smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = sl.pdl ? 1 : 0;
    void* kargs[] = {const_cast<char*>(sl.args.data())};
    return cudaLaunchKernelExC(&cfg, sl.func, kargs) == cudaSuccess ? HGNN_OK : HGNN_ERR_CUDA;
}

// captures the recorded chain on a private stream and remembers the kernel node of every slot
bool replay_build(ReplayGraph& g, const hgnn_eng_recorder_t& rec) {
    replay_destroy(g);
    static thread_local cudaStream_t cap = nullptr;
    if (!cap && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    bool ok = true;
    for (const hgnn_eng_slot_t& sl : rec.slots) {
        if (replay_launch_slot(sl, cap) != HGNN_OK) { ok = false; break; }
        cudaStreamCaptureStatus st;
        unsigned long long id = 0;
        cudaGraph_t cg = nullptr;
        const cudaGraphNode_t* deps = nullptr;
        size_t nd = 0;
        if (cudaStreamGetCaptureInfo(cap, &st, &id, &cg, &deps, &nd) != cudaSuccess || nd != 1) { ok = false; break; }
        g.nodes.push_back(deps[0]);            // the node just added: what the next launch would depend on
        g.funcs.push_back(sl.func);
        g.block.push_back(sl.block);
        g.pdl.push_back(sl.pdl);
        g.smem.push_back(sl.smem);
    }
    cudaGraph_t graph = nullptr;
    if (cudaStreamEndCapture(cap, &graph) != cudaSuccess || !graph) ok = false;
    g.graph = graph;
    if (ok && cudaGraphInstantiate(&g.exec, g.graph, 0) != cudaSuccess) ok = false;
    if (!ok) { cudaGetLastError(); replay_destroy(g); }
    return ok;
}

// writes the recorded slots into the graph (rebuilding it when the kernel sequence changed) and launches it
int replay_run(ReplayGraph& g, const hgnn_eng_recorder_t& rec, cudaStream_t s) {
    const size_t n = rec.slots.size();
    if (n == 0) return HGNN_OK;
    bool same = g.exec != nullptr && g.nodes.size() == n;
    for (size_t i = 0; same && i < n; ++i) {
        const hgnn_eng_slot_t& sl = rec.slots[i];
        same = g.funcs[i] == sl.func && g.block[i] == sl.block && g.pdl[i] == sl.pdl && g.smem[i] == sl.smem;
    }
    if (!same) {
        if (g.failures > 8 || !replay_build(g, rec)) { ++g.failures; return HGNN_ERR_CUDA; }
    } else {
        for (size_t i = 0; i < n; ++i) {
            const hgnn_eng_slot_t& sl = rec.slots[i];
            void* kargs[] = {const_cast<char*>(sl.args.data())};
            cudaKernelNodeParams kp = {};
            kp.func = const_cast<void*>(sl.func);
            kp.gridDim = dim3(sl.grid);
            kp.blockDim = dim3(sl.block);
            kp.sharedMemBytes = sl.smem;
            kp.kernelParams = kargs;
            if (cudaGraphExecKernelNodeSetParams(g.exec, g.nodes[i], &kp) != cudaSuccess) {
                cudaGetLastError();
                ++g.failures;
                replay_destroy(g);
                return HGNN_ERR_CUDA;
            }
        }
    }
    if (cudaGraphLaunch(g.exec, s) != cudaSuccess) {
        hgnn_set_error("hgnn_program: cudaGraphLaunch: %s", cudaGetErrorString(cudaGetLastError()));
        return HGNN_ERR_CUDA;
    }
    return HGNN_OK;
}

// one graph per (program, direction) and thread
ReplayGraph& replay_graph_of(const hgnn_program_t* prog, int direction) {
    static thread_local std::map<std::pair<const void*, int>, ReplayGraph> graphs;
    return graphs[std::make_pair(static_cast<const void*>(prog), direction)];
}

#define PROG_CALL(expr)                    \
    do {                                   \
        int rc_ = (expr);                  \
        if (rc_ != HGNN_OK) return rc_;    \
        g_program_launches.fetch_add(1);   \
    } while (0)

// rc != NULL: a launch recorder is installed - the side kernels are recorded instead of launched, and the direct
// launches that must follow them are deferred into rc->post
static int program_fwd_impl(const hgnn_program_t* prog, const hgnn_batch_t* b, const float* X, const float* XL,
                            const long long* addr, float* work, double* arena, float* running, float* out,
                            hgnn_stream_t stream, ReplayCtx* rc) {
    WorkLayout w;
    HGNN_REQUIRE(prog && b && X && addr && work && arena && out, "null argument");
    HGNN_REQUIRE(plan_work(prog, b->Rn, b->Rm, &w) && check_program(prog, b), "malformed program or batch");
    HGNN_REQUIRE(!prog->dual || XL, "line-graph model without XL");
    cudaStream_t s = to_stream(stream);
    if (prog->arena_doubles > 0) {
        cudaError_t e = cudaMemsetAsync(arena, 0, (size_t)prog->arena_doubles * sizeof(double), s);
        if (e != cudaSuccess) {
            hgnn_set_error("hgnn_program_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
            return HGNN_ERR_CUDA;
        }
        g_program_launches.fetch_add(1);
    }
    int m0 = 0, m1 = 0, fwd_expand = -1, bwd_expand_unused = -1;
    mega_decide(prog, b, &m0, &m1, &fwd_expand, &bwd_expand_unused);
    const bool collapse = use_collapse(prog, b);      // per-side kernels on the collapsed line graph
    if (collapse) fwd_expand = -1;                    // nobody reads the skipped rows then
    hgnn_op_t edge_F_collapsed[3];
    if (collapse) {
        // Same operators.  The entry-count hint (it picks the gather batch size) deliberately stays the one of the full
        // operator: with the hint of the active rows (1.5 entries per row) the forward takes the <4, 4> variant, which
        // needs 131 registers (3 CTAs per SM) or spills at 128, and measured 9.7 us against 7.2 us for <8, 4> on 444 CTAs
        // (profiles/logs/bench_r2r_edge_fwd_small_batch.log).
        for (int k = 0; k < 3; ++k) edge_F_collapsed[k] = b->edge_ops[k];
    }
    for (int i = 0; i < prog->n_sides; ++i) {
        if (m1 > m0 && i == m0) {      // sides [m0, m1): one persistent kernel (mega.cu)
            mk::Params P;
            mega_params(prog, b, w, X, XL, addr, work, nullptr, nullptr, arena, m0, m1, &P);
            P.expand = fwd_expand;
            PROG_CALL(hgnn_mega_launch_fwd(P, s));
            i = m1 - 1;
            continue;
        }
        const hgnn_prog_side_t& sd = prog->sides[i];
        const bool node = sd.kind == 0;
        hgnn_side_t st;
        st.R = node ? b->Rn : b->Rm;
        st.ops = node ? b->node_ops : (collapse ? edge_F_collapsed : b->edge_ops);
        st.n_ops = b->n_ops;
        st.Xs = tensor_ptr(prog, w, sd.src_self, X, XL, work);
        st.Fs = prog->tensors[sd.src_self].F;
        hgnn_bn_ref_t bs_ = bn_ref(prog, b, sd.src_self, addr, arena), bc_;
        const bool cross = sd.src_cross >= 0;
        if (cross) {
            st.p_rowptr = node ? b->p_rowptr : b->pt_rowptr;
            st.p_col = node ? b->p_col : b->pt_col;
            st.p_pm = node ? b->p_pm : b->pt_pm;
            st.p_pd = node ? b->p_pd : b->pt_pd;
            st.Xc = tensor_ptr(prog, w, sd.src_cross, X, XL, work);
            st.Fc = prog->tensors[sd.src_cross].F;
            st.p_nnz = b->p_nnz;
            bc_ = bn_ref(prog, b, sd.src_cross, addr, arena);
        } else {
            st.p_rowptr = st.p_col = nullptr;
            st.p_pm = st.p_pd = st.Xc = nullptr;
            st.Fc = 0;
            st.p_nnz = 0;
        }
        st.roww = (!node && collapse) ? b->ew : nullptr;
        st.rowmap = (!node && collapse) ? b->erow : nullptr;
        if (st.rowmap) st.R = b->n_act;                      // only the active line-graph rows
        const bool readout = sd.out < 0;
        float* Z = work + (readout ? w.readout_off : w.off[sd.out]);
        double* acc_out = readout ? nullptr : arena + prog->tensors[sd.out].acc_f;
        hgnn_eng_set_pdl(i >= 1 && i != m1);      // the stream predecessor is the previous side's forward kernel
        const int rc_fwd = hgnn_lg_side_fwd(&st, &bs_, cross ? &bc_ : nullptr, param(addr, sd.Wa), param(addr, sd.ba),
                                            sd.Ha, param(addr, sd.Wb), param(addr, sd.bb), sd.Hb, sd.relu_from, Z,
                                            acc_out, nullptr, stream);
        hgnn_eng_set_pdl(false);
        PROG_CALL(rc_fwd);
        if (readout) {  // sum over all Nmax slots; padded slots add fc.bias (layers_mnb.py:92, :386)
            const int Fo = sd.Ha + sd.Hb;
            const float* bias = param(addr, sd.ba);
            auto fn = [=]() { return hgnn_segment_sum(Z, b->bs, Fo, b->node_off, b->pad_n, bias, out, stream); };
            if (rc) rc->post.push_back(fn);
            else PROG_CALL(fn());
        }
    }
    if (running && prog->n_bn > 0) {
        auto fn = [=]() {
            return hgnn_bn_running_update_k(arena, prog->bn_acc_off, prog->bn_F, prog->bn_rows_kind, b->Rn, b->Rm,
                                            prog->bn_run_off, prog->n_bn, prog->momentum, running, stream);
        };
        if (rc) rc->post.push_back(fn);
        else PROG_CALL(fn());
    }
    return HGNN_OK;
}

// after a recording pass: the recorded kernels as one graph launch (or, should the graph be unavailable, one by one),
// then the deferred direct launches
static int replay_finish(const hgnn_program_t* prog, int direction, ReplayCtx& rc, cudaStream_t s) {
    int r = replay_run(replay_graph_of(prog, direction), rc.rec, s);
    if (r != HGNN_OK) {
        cudaGetLastError();
        for (const hgnn_eng_slot_t& sl : rc.rec.slots)
            if (replay_launch_slot(sl, s) != HGNN_OK) {
                hgnn_set_error("hgnn_program: launch of a recorded kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
                return HGNN_ERR_CUDA;
            }
    }
    for (auto& f : rc.post) PROG_CALL(f());
    return HGNN_OK;
}

extern "C" int hgnn_program_fwd(const hgnn_program_t* prog, const hgnn_batch_t* b, const float* X, const float* XL,
                                const long long* addr, float* work, double* arena, float* running, float* out,
                                hgnn_stream_t stream) {
    cudaStream_t s = to_stream(stream);
    if (prog && b && prog->sides && prog->tensors && replay_wanted(prog, b, s)) {
        ReplayCtx rc;
        hgnn_eng_set_recorder(&rc.rec);
        const int r = program_fwd_impl(prog, b, X, XL, addr, work, arena, running, out, stream, &rc);
        hgnn_eng_set_recorder(nullptr);
        if (r == HGNN_OK) return replay_finish(prog, 0, rc, s);
        // the recording pass launched nothing but the (idempotent) arena memset: issue the pass directly
    }
    return program_fwd_impl(prog, b, X, XL, addr, work, arena, running, out, stream, nullptr);
}

// one scratch region per edge side whose transposed operator has a run-length part
static long long side_rng_bytes(const hgnn_program_t* prog, const hgnn_batch_t* b) {
    if (!prog->dual || !b->edge_ops_T || b->n_ops < 3) return 0;
    const hgnn_op_t& o = b->edge_ops_T[2];
    if (o.kind != HGNN_OP_CSR || !o.rng_rowptr) return 0;
    return hgnn_lg_rng_scratch_bytes(o.rng_n);
}

extern "C" long long hgnn_program_rng_scratch_bytes(const hgnn_program_t* prog, const hgnn_batch_t* b) {
    if (!prog || !b || !prog->sides) return 0;
    long long n = 0;
    for (int i = 0; i < prog->n_sides; ++i) n += prog->sides[i].kind == 1 ? 1 : 0;
    return side_rng_bytes(prog, b) * n;
}

static int program_bwd_impl(const hgnn_program_t* prog, const hgnn_batch_t* b, const float* X, const float* XL,
                            const long long* addr, const float* work, float* gwork, double* arena,
                            const float* g_out, float* gX, float* gflat, void* rng_scratch,
                            long long rng_scratch_bytes, hgnn_stream_t stream, ReplayCtx* rc) {
    WorkLayout w;
    HGNN_REQUIRE(prog && b && X && addr && work && gwork && arena && g_out && gflat, "null argument");
    HGNN_REQUIRE(plan_work(prog, b->Rn, b->Rm, &w) && check_program(prog, b), "malformed program or batch");
    cudaStream_t s = to_stream(stream);
    std::vector<char> started(prog->n_tensors, 0);
    const long long rng_per_side = side_rng_bytes(prog, b), rng_total = hgnn_program_rng_scratch_bytes(prog, b);
    long long rng_used = 0;
    if (rng_scratch && rng_per_side > 0 && rng_scratch_bytes >= rng_total) {
        if (cudaMemsetAsync(rng_scratch, 0, (size_t)rng_total, s) != cudaSuccess) {
            hgnn_set_error("hgnn_program_bwd: cudaMemsetAsync(rng_scratch) failed");
            return HGNN_ERR_CUDA;
        }
        g_program_launches.fetch_add(1);
    } else {
        rng_scratch = nullptr;
    }
    auto grad_ptr = [&](int t) -> float* { return t == 0 ? gX : gwork + w.off[t]; };
    auto wants_grad = [&](int t) { return prog->tensors[t].bn_weight >= 0 || (t == 0 && gX != nullptr); };
    // ---- pre-pass in backward order: who writes which gradient first (the later writers accumulate)
    struct Flags { char need_self, acc_self, need_cross, acc_cross, out_started; };
    std::vector<Flags> fl(prog->n_sides);
    for (int i = prog->n_sides - 1; i >= 0; --i) {
        const hgnn_prog_side_t& sd = prog->sides[i];
        Flags& f = fl[i];
        f.out_started = sd.out >= 0 ? started[sd.out] : 1;
        f.need_self = wants_grad(sd.src_self) ? 1 : 0;
        f.acc_self = started[sd.src_self];
        if (f.need_self) started[sd.src_self] = 1;
        f.need_cross = f.acc_cross = 0;
        if (sd.src_cross >= 0) {
            f.need_cross = wants_grad(sd.src_cross) ? 1 : 0;
            f.acc_cross = started[sd.src_cross];
            if (f.need_cross) started[sd.src_cross] = 1;
        }
    }
    if (rc)      // a zero fill between recorded kernels cannot be deferred: such programs are issued directly
        for (int i = 0; i < prog->n_sides; ++i)
            if (prog->sides[i].out >= 0 && !fl[i].out_started) {
                hgnn_set_error("hgnn_program_bwd: recording: side %d has an unused output", i);
                return HGNN_ERR_ARG;
            }
    int m0 = 0, m1 = 0, fwd_expand_unused = -1, bwd_expand = -1;
    mega_decide(prog, b, &m0, &m1, &fwd_expand_unused, &bwd_expand);     // the same decision as the forward
    const bool collapse = use_collapse(prog, b);
    if (collapse) bwd_expand = -1;
    hgnn_op_t edge_T_collapsed[3];
    if (collapse) {      // [I, D, btc]: plain CSR, the multiplicity of a representative folded into its entries
        for (int k = 0; k < 3; ++k) edge_T_collapsed[k] = b->edge_ops_T[k];
        hgnn_op_t& o = edge_T_collapsed[2];
        o.rowptr = b->btc_rowptr; o.col = b->btc_col; o.val = b->btc_val;
        o.rng_rowptr = nullptr; o.rng_id = nullptr; o.rng_val = nullptr; o.rng_lo = nullptr; o.rng_hi = nullptr;
        o.rng_n = 0;
        o.nnz = b->btc_nnz;
    }
    if (rc) rc->restartable = false;      // from here on direct launches add into the arena
    for (int i = prog->n_sides - 1; i >= 0; --i) {
        if (m1 > m0 && i == m1 - 1) {     // sides [m0, m1) in reverse: one persistent kernel (mega.cu)
            mk::Params P;
            mega_params(prog, b, w, X, XL, addr, work, gwork, gX, arena, m0, m1, &P);
            for (int k = m0; k < m1; ++k) {
                mk::Side& o = P.s[k - m0];
                o.need_self = fl[k].need_self; o.acc_self = fl[k].acc_self;
                o.need_cross = fl[k].need_cross; o.acc_cross = fl[k].acc_cross;
            }
            P.expand = bwd_expand;
            PROG_CALL(hgnn_mega_launch_bwd(P, s));
            i = m0;
            continue;
        }
        const hgnn_prog_side_t& sd = prog->sides[i];
        const bool node = sd.kind == 0;
        hgnn_side_bwd_t d;
        const int Fout = sd.Ha + sd.Hb;
        if (sd.out < 0) {   // readout: gPre = g_out broadcast over the rows of each graph
            float* G = gwork + w.readout_off;
            PROG_CALL(hgnn_readout_bwd_prep(g_out, b->bs, Fout, b->node_off, b->pad_n, G, arena + sd.db_off, stream));
            d.gY = G;
            d.Z = nullptr;
            d.acc_f = d.acc_b = nullptr;
            d.bn_weight = nullptr;
            d.Rg = b->Rn;
        } else {
            const hgnn_prog_tensor_t& T = prog->tensors[sd.out];
            const size_t n = (size_t)rows_of(prog, b, sd.out) * T.F;
            if (!fl[i].out_started) {   // output never used downstream: zero gradient
                if (n && cudaMemsetAsync(gwork + w.off[sd.out], 0, n * sizeof(float), s) != cudaSuccess) {
                    hgnn_set_error("hgnn_program_bwd: cudaMemsetAsync failed");
                    return HGNN_ERR_CUDA;
                }
                g_program_launches.fetch_add(1);
            }
            d.gY = gwork + w.off[sd.out];
            d.Z = work + w.off[sd.out];
            d.acc_f = arena + T.acc_f;
            d.acc_b = arena + T.acc_b;
            d.bn_weight = param(addr, T.bn_weight);
            d.Rg = rows_of(prog, b, sd.out);
        }
        d.Fg = Fout;
        d.relu_from = sd.relu_from;
        d.Wa = param(addr, sd.Wa);
        d.Ha = sd.Ha;
        d.Wb = param(addr, sd.Wb);
        d.Hb = sd.Hb;
        const int Fs = prog->tensors[sd.src_self].F;
        const int Fc = sd.src_cross >= 0 ? prog->tensors[sd.src_cross].F : 0;
        d.Cin = b->n_ops * Fs + 2 * Fc;
        d.dW_bins = arena + sd.dW_off;
        d.db_bins = arena + sd.db_off;
        // self part
        d.R_self = node ? b->Rn : b->Rm;
        d.ops_T = node ? b->node_ops_T : (collapse ? edge_T_collapsed : b->edge_ops_T);
        d.roww_self = (!node && collapse) ? b->ew : nullptr;
        d.roww_cross = nullptr;
        d.rowmap_self = d.roww_self ? b->erow : nullptr;
        d.rowmap_cross = nullptr;
        if (d.rowmap_self) d.R_self = b->n_act;
        d.n_ops = b->n_ops;
        d.Xs = tensor_ptr(prog, w, sd.src_self, X, XL, work);
        d.Fs = Fs;
        d.bn_self = bn_ref(prog, b, sd.src_self, addr, arena);
        d.gXs = fl[i].need_self ? grad_ptr(sd.src_self) : nullptr;
        d.accumulate_self = fl[i].acc_self;
        d.acc_b_self = prog->tensors[sd.src_self].bn_weight >= 0 ? arena + prog->tensors[sd.src_self].acc_b : nullptr;
        // cross part
        d.R_cross = 0;
        d.pt_rowptr = d.pt_col = nullptr;
        d.pt_pm = d.pt_pd = d.Xc = nullptr;
        d.Fc = 0;
        d.gXc = nullptr;
        d.accumulate_cross = 0;
        d.acc_b_cross = nullptr;
        d.pt_nnz = 0;
        d.bn_cross = bn_ref(prog, b, 0, addr, arena);   // tensor 0 is never normalised: an empty reference
        if (sd.src_cross >= 0) {
            d.R_cross = rows_of(prog, b, sd.src_cross);
            // rows = the cross tensor's rows: the other side's incidence pattern
            d.pt_rowptr = node ? b->pt_rowptr : b->p_rowptr;
            d.pt_col = node ? b->pt_col : b->p_col;
            d.pt_pm = node ? b->pt_pm : b->p_pm;
            d.pt_pd = node ? b->pt_pd : b->p_pd;
            d.Xc = tensor_ptr(prog, w, sd.src_cross, X, XL, work);
            d.Fc = Fc;
            d.pt_nnz = b->p_nnz;
            d.bn_cross = bn_ref(prog, b, sd.src_cross, addr, arena);
            d.gXc = fl[i].need_cross ? grad_ptr(sd.src_cross) : nullptr;
            d.accumulate_cross = fl[i].acc_cross;
            d.acc_b_cross =
                prog->tensors[sd.src_cross].bn_weight >= 0 ? arena + prog->tensors[sd.src_cross].acc_b : nullptr;
            d.roww_cross = (node && collapse) ? b->ew : nullptr;      // the cross rows of a node side are line-graph rows
            d.rowmap_cross = d.roww_cross ? b->erow : nullptr;
            if (d.rowmap_cross) d.R_cross = b->n_act;
        }
        d.skip_dw = 0;
        d.rng_scratch = nullptr;
        if (rng_scratch && !node && !collapse) {
            d.rng_scratch = static_cast<char*>(rng_scratch) + rng_used;
            rng_used += rng_per_side;
        }
        hgnn_eng_set_pdl(i < prog->n_sides - 1 && i != m0 - 1);   // the stream predecessor is the backward kernel of side i + 1
        const int rc_bwd = hgnn_lg_side_bwd(&d, stream);
        hgnn_eng_set_pdl(false);
        PROG_CALL(rc_bwd);
    }
    {
        auto fn = [=]() {
            return hgnn_bins_reduce(arena, prog->red_off, prog->red_nb, prog->red_stride, prog->red_cnt, prog->n_flat,
                                    gflat, stream);
        };
        if (rc) rc->post.push_back(fn);
        else PROG_CALL(fn());
    }
    return HGNN_OK;
}

extern "C" int hgnn_program_bwd(const hgnn_program_t* prog, const hgnn_batch_t* b, const float* X, const float* XL,
                                const long long* addr, const float* work, float* gwork, double* arena,
                                const float* g_out, float* gX, float* gflat, void* rng_scratch,
                                long long rng_scratch_bytes, hgnn_stream_t stream) {
    cudaStream_t s = to_stream(stream);
    // the pre-pass of the implementation refuses (before launching anything) programs it cannot record
    if (prog && b && prog->sides && prog->tensors && replay_wanted(prog, b, s)) {
        ReplayCtx rc;
        hgnn_eng_set_recorder(&rc.rec);
        const int r = program_bwd_impl(prog, b, X, XL, addr, work, gwork, arena, g_out, gX, gflat, rng_scratch,
                                       rng_scratch_bytes, stream, &rc);
        hgnn_eng_set_recorder(nullptr);
        if (r == HGNN_OK) return replay_finish(prog, 1, rc, s);
        if (!rc.restartable) return r;     // failed after a non-idempotent direct launch
    }
    return program_bwd_impl(prog, b, X, XL, addr, work, gwork, arena, g_out, gX, gflat, rng_scratch, rng_scratch_bytes,
                            stream, nullptr);
}
