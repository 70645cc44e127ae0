// engine_row4.cuh -- thread-per-row engine kernels for the script-default width (h = 2: every
// state is 4 floats = one float4 per node / line-graph node).  Included by engine.cu inside
// namespace eng.
//
// At this width the generic tile kernels are instruction- and barrier-bound (profiles/README.md:
// ~1 600 thread-instructions and four CTA barriers per row).  Here one thread owns one row from
// the first load to the last store: all operator blocks, the concatenated x1 vector (20 floats),
// the 20x4 mat-vec and - in the backward - the 48+32 dW accumulators live in registers; there is
// no shared-memory tile and no barrier in the row loop.  Loads are issued in three dependent
// rounds (row pointers -> entries -> feature rows), batches of up to 4 entries in flight.
#pragma once
#include "engine_mma.cuh"

#define R4_THREADS 128

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// profiling aid (HGNN_B200_ABLATE bit 16): timeline of the traced launches of a step, one slot per launch in issue
// order: [min CTA start, min "wait passed", max "wait passed", max CTA end] (ns, %globaltimer); read and reset with
// hgnn_debug_ktrace.  Works inside a replayed CUDA graph (the slot is a kernel argument).
// per-CTA stamps of the last traced width-4 launch (HGNN_B200_ABLATE bit 8): hgnn_debug_cta_times / _phases.
// backward: times = (start, end, is_self), phase = (end of row loop, end of range phase, coefficients ready);
// forward:  times = (start, end, 2),       phase = (end of row loop, producer wait passed, coefficients ready)
__device__ unsigned long long g_cta_times[2048 * 3];
__device__ unsigned long long g_cta_phase[2048 * 3];
#define KTRACE_SLOTS 1024
__device__ unsigned long long g_ktrace[KTRACE_SLOTS * 4];
__device__ __forceinline__ void ktrace_start(int slot) {
    if (slot >= 0 && threadIdx.x == 0) atomicMin(&g_ktrace[slot * 4], global_ns());
}
__device__ __forceinline__ void ktrace_waited(int slot) {
    if (slot >= 0 && threadIdx.x == 0) {
        const unsigned long long t = global_ns();
        atomicMin(&g_ktrace[slot * 4 + 1], t);
        atomicMax(&g_ktrace[slot * 4 + 2], t);
    }
}
__device__ __forceinline__ void ktrace_end(int slot) {
    if (slot >= 0 && threadIdx.x == 0) atomicMax(&g_ktrace[slot * 4 + 3], global_ns());
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 acc) {
    acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
    return acc;
}
__device__ __forceinline__ float4 f4_affine(float4 x, float4 s, float4 t) {
    return make_float4(fmaf(x.x, s.x, t.x), fmaf(x.y, s.y, t.y), fmaf(x.z, s.z, t.z), fmaf(x.w, s.w, t.w));
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// sum_k val[k] * X[col[k]] (raw rows) and sum_k val[k]; batches of 4 in flight
__device__ __forceinline__ void csr_gather4(const int* __restrict__ col, const float* __restrict__ val, int k0,
                                            int k1, const float* __restrict__ X, float4& acc, float& wsum) {
    for (int k = k0; k < k1; k += 4) {
        int c[4];
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool on = k + j < k1;
            c[j] = __ldg(col + (on ? k + j : k));
            v[j] = on ? __ldg(val + k + j) : 0.f;
        }
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ld4(X + (size_t)c[j] * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc = f4_fma(v[j], x[j], acc);
            wsum += v[j];
        }
    }
}

// ---- warp-level batch-norm prologue for width-4 tensors: the binned (sum, sum^2) accumulators are
// hgnn_ws_bins(8) x 8 doubles = four independent loads per lane (hgnn_bins8_lane); two shuffles fold the bins, eight more
// broadcast the totals, and every thread derives the vectors in registers - no shared memory and
// no block barrier.
__device__ __forceinline__ void warp_totals8(const double* __restrict__ acc, double tot[8]) {
    double v = hgnn_bins8_lane(acc);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
#pragma unroll
    for (int c = 0; c < 8; ++c) tot[c] = __shfl_sync(0xffffffffu, v, c);
}

struct Bn4 {
    float4 sc, sh, mu, rs;
    bool on;
};

// One warp, one feature per lane (f = lane & 3): the folded totals of a width-4 tensor's (sum, sum^2) block -> this
// lane's scale / shift / mean / 1/std.  `v` is the lane's part of the block (hgnn_bins8_lane).  Lanes 0..3 hold the four
// features (the other lanes repeat them).  ~25 instructions per lane on the critical path instead of the 4-feature loop
// behind eight broadcast shuffles that bn4_from_totals runs.
__device__ __forceinline__ void bn4_lane(double v, float w, float b, double inv_n, float& sc, float& sh, float& mu, float& rs) {
    const int f = threadIdx.x & 3;
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    const double sum = __shfl_sync(0xffffffffu, v, f), sq = __shfl_sync(0xffffffffu, v, 4 + f);
    const double m = sum * inv_n;
    const double var = fma(-m, m, sq * inv_n);
    rs = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
    mu = (float)m;
    sc = w * rs;
    sh = b - w * mu * rs;
}

__device__ __forceinline__ Bn4 bn4_from_ref(const BnRef& r) {
    Bn4 o;
    o.on = true;
    if (r.affine) {
        o.sc = ld4(r.affine);
        o.sh = ld4(r.affine + 4);
        o.mu = f4_zero();
        o.rs = make_float4(1.f, 1.f, 1.f, 1.f);
        return o;
    }
    if (!r.acc) {
        o.on = false;
        o.sc = make_float4(1.f, 1.f, 1.f, 1.f);
        o.sh = f4_zero();
        o.mu = f4_zero();
        o.rs = make_float4(1.f, 1.f, 1.f, 1.f);
        return o;
    }
    double tot[8];
    warp_totals8(r.acc, tot);
    // E[x^2] - mean^2 in fp64 (the only cancellation-prone step); everything after in fp32 - fp64
    // divide / sqrt are long software sequences and every thread runs this prologue
    const float w = r.w[0], b = r.b[0];
    const double inv_n = r.inv_n;
    float sc[4], sh[4], mu[4], rs[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const double m = tot[f] * inv_n;
        const double var = fma(-m, m, tot[4 + f] * inv_n);
        const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
        mu[f] = (float)m;
        rs[f] = r_;
        sc[f] = w * r_;
        sh[f] = b - w * mu[f] * r_;
    }
    o.sc = make_float4(sc[0], sc[1], sc[2], sc[3]);
    o.sh = make_float4(sh[0], sh[1], sh[2], sh[3]);
    o.mu = make_float4(mu[0], mu[1], mu[2], mu[3]);
    o.rs = make_float4(rs[0], rs[1], rs[2], rs[3]);
    return o;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct Fwd4Args {
    int R;
    int n_csr;                       // K - 2 CSR operators after [IDENT, DIAG]
    const float* diag;
    const int* rowptr[2]; const int* col[2]; const float* val[2];
    const float* Xs; BnRef bn_s;
    const int* p_rowptr; const int* p_col; const float* p_pm; const float* p_pd;   // NULL: no cross part
    const float* Xc; BnRef bn_c;
    const float* Wa; const float* ba; int Ha;
    const float* Wb; const float* bb; int Hb;
    int relu_from, Cin;
    float* Z;
    double* acc_out;
    float* X1;                       // optional: the concatenated x1 rows (R, Cin) for the dW pass
    const float* roww;               // optional per-row weight (collapsed line graph): rows with weight <= 0 are skipped,
                                     // the others enter the batch-norm sums weight times
    const int* rowmap;               // optional list of the R rows to compute (the active rows); NULL: rows 0..R-1
    int ablate;                      // timing experiments only (HGNN_B200_ABLATE): 1 no gathers, 2 no BN, 4 no epilogue
    int trace_slot;                  // >= 0: record this launch in g_ktrace (HGNN_B200_ABLATE bit 16)
};

// One batch of a CSR gather, split in two so that the index loads of several operators can be in
// flight together before any feature row is requested: `Batch::load_entries` (columns / values of
// up to B entries, predicated), then `load_rows`, then `accumulate`.
template <int B, bool TWO>
struct GatherBatch {
    int c[B];
    float v[B], v2[TWO ? B : 1];
    float4 x[B];
    __device__ __forceinline__ void load_entries(const int* __restrict__ col, const float* __restrict__ val,
                                                 const float* __restrict__ val2, int k, int k1) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const bool on = k + j < k1;
            c[j] = on ? __ldg(col + k + j) : -1;
            v[j] = on ? __ldg(val + k + j) : 0.f;
            if (TWO) v2[j] = on ? __ldg(val2 + k + j) : 0.f;
        }
    }
    __device__ __forceinline__ void load_rows(const float* __restrict__ X) {
#pragma unroll
        for (int j = 0; j < B; ++j) x[j] = c[j] >= 0 ? ld4(X + (size_t)c[j] * 4) : f4_zero();
    }
    __device__ __forceinline__ void accumulate(float4& acc, float& ws, float4& acc2, float& ws2) const {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            acc = f4_fma(v[j], x[j], acc);
            ws += v[j];
            if (TWO) {
                acc2 = f4_fma(v2[j], x[j], acc2);
                ws2 += v2[j];
            }
        }
    }
};

// Programmatic dependent launch (PDL): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may start while its producer is still running;
// everything before pdl_wait() may only touch data the producer does not write (parameters, graph
// structure).  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// (sum, sum^2)-style totals of a width-4 tensor from its 32 binned doubles, given the lane's own load
__device__ __forceinline__ void warp_totals8_from(double v, double tot[8]) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
#pragma unroll
    for (int c = 0; c < 8; ++c) tot[c] = __shfl_sync(0xffffffffu, v, c);
}

// batch-norm vectors from the totals; w, b already in registers; inv_n computed on the host
__device__ __forceinline__ Bn4 bn4_from_totals(const double tot[8], float w, float b, double inv_n) {
    Bn4 o;
    o.on = true;
    float sc[4], sh[4], mu[4], rs[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const double m = tot[f] * inv_n;
        const double var = fma(-m, m, tot[4 + f] * inv_n);
        const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
        mu[f] = (float)m;
        rs[f] = r_;
        sc[f] = w * r_;
        sh[f] = b - w * mu[f] * r_;
    }
    o.sc = make_float4(sc[0], sc[1], sc[2], sc[3]);
    o.sh = make_float4(sh[0], sh[1], sh[2], sh[3]);
    o.mu = make_float4(mu[0], mu[1], mu[2], mu[3]);
    o.rs = make_float4(rs[0], rs[1], rs[2], rs[3]);
    return o;
}

// Issue-then-resolve form of bn4_from_ref: `issue` puts the loads in flight (accumulators after
// pdl_wait, the scalar affine before it), `resolve` does the arithmetic.
struct Bn4Loader {
    double v;
    float w, b;
    int mode;       // 0: identity, 1: precomputed affine, 2: batch statistics
    __device__ __forceinline__ void issue_params(const BnRef& r) {
        mode = r.affine ? 1 : (r.acc ? 2 : 0);
        w = b = 0.f;
        if (mode == 2) { w = __ldg(r.w); b = __ldg(r.b); }
    }
    __device__ __forceinline__ void issue_acc(const BnRef& r) {
        v = 0.0;
        if (mode == 2) v = hgnn_bins8_lane(r.acc);
    }
    __device__ __forceinline__ Bn4 resolve(const BnRef& r) const {
        if (mode != 2) return bn4_from_ref(r);
        double tot[8];
        warp_totals8_from(v, tot);
        return bn4_from_totals(tot, w, b, r.inv_n);
    }
};

// Sum of NV per-lane values over the warp with a reduce-scatter butterfly: at every level a lane
// keeps half of its values and ships the other half, so NV values cost ~NV shuffles instead of
// 5 NV.  NV is a power of two.  Afterwards lane l holds the warp totals of the original indices
//   NV = 64: 2l, 2l+1 in val[0], val[1];  NV = 32: l in val[0];  NV = 16: l >> 1;  NV = 8: l >> 2.
template <typename T, int NV>
__device__ __forceinline__ void warp_reduce_scatter(T (&val)[NV]) {
    const int lane = threadIdx.x & 31;
    int n = NV;
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
        const int mask = 16 >> bit;            // partner distance 16, 8, 4, 2, 1
        const bool upper = (lane & mask) != 0;
        if (n > 1) {
            const int half = n >> 1;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < half) {
                    // lower lanes keep [0, half), upper lanes keep [half, n)
                    const T send = upper ? val[i] : val[i + half];
                    const T keep = upper ? val[i + half] : val[i];
                    val[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
                }
            }
            n = half;
        } else {
            val[0] += __shfl_xor_sync(0xffffffffu, val[0], mask);
        }
    }
}

// BA / BP: entries per batch of the (first) self operator / of the incidence pair.  The host picks
// them from the average row length, so that a typical row needs ONE batch: all its index loads go
// out in one round and all its feature-row loads in the next (three dependent rounds per row in
// total: row pointers -> entries -> rows), instead of two rounds per batch of four.
//
// Phases: (0) parameters and graph structure only - weights to shared memory, the row pointers and
// the first batch of entries of the thread's first row; under PDL this overlaps the producer's tail.
// (1) after pdl_wait: batch-norm accumulators and feature rows, all issued before the first use.
template <int NCSR, bool CROSS, int BA, int BP>
__global__ void __launch_bounds__(R4_THREADS)
fwd_row4_kernel(const Fwd4Args a) {
    constexpr int NB = 2 + NCSR + (CROSS ? 2 : 0);     // float4 blocks of x1
    __shared__ __align__(16) float W[4 * NB * 4];       // [o][Cin]
    __shared__ __align__(16) float bias[4];
    __shared__ double red[(R4_THREADS / 32) * 8];
    __shared__ __align__(16) float bnv[16];            // self: scale, shift | cross: scale, shift (one warp each computes them)
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    ktrace_start(a.trace_slot);
    const bool stamp = (a.ablate & 8) && tid == 0 && blockIdx.x < 2048;
    if (stamp) { g_cta_times[blockIdx.x * 3] = global_ns(); g_cta_times[blockIdx.x * 3 + 2] = 2; }
    // ---- phase 0
    for (int i = tid; i < 4 * NB * 4; i += R4_THREADS) {
        const int o = i / (NB * 4), c = i - o * (NB * 4);
        W[i] = (o < a.Ha) ? a.Wa[(size_t)o * a.Cin + c] : a.Wb[(size_t)(o - a.Ha) * a.Cin + c];
    }
    if (tid < 4) bias[tid] = (tid < a.Ha) ? (a.ba ? a.ba[tid] : 0.f) : (a.bb ? a.bb[tid - a.Ha] : 0.f);
    Bn4Loader ls, lc;
    ls.issue_params(a.bn_s);
    if (CROSS) lc.issue_params(a.bn_c);
    const int stride = gridDim.x * R4_THREADS;
    int row = blockIdx.x * R4_THREADS + tid;
    float d = 0.f, rw = 1.f;
    int rr = 0;                                        // the row behind list entry `row`
    int k0[NCSR > 0 ? NCSR : 1], k1[NCSR > 0 ? NCSR : 1], p0 = 0, p1 = 0;
    GatherBatch<BA, false> ga;
    GatherBatch<BP, true> gb;
    auto load_structure = [&](int idx) {
        const int r = a.rowmap ? __ldg(a.rowmap + idx) : idx;
        rr = r;
        d = __ldg(a.diag + r);
#pragma unroll
        for (int t = 0; t < NCSR; ++t) {
            k0[t] = __ldg(a.rowptr[t] + r);
            k1[t] = __ldg(a.rowptr[t] + r + 1);
        }
        if (CROSS) {
            p0 = __ldg(a.p_rowptr + r);
            p1 = __ldg(a.p_rowptr + r + 1);
        }
        rw = a.roww ? __ldg(a.roww + r) : 1.f;
        if ((a.ablate & 1) || rw <= 0.f) { k1[0] = k0[0]; p1 = p0; }
        if (NCSR > 0) ga.load_entries(a.col[0], a.val[0], nullptr, k0[0], k1[0]);
        if (CROSS) gb.load_entries(a.p_col, a.p_pm, a.p_pd, p0, p1);
    };
    if (row < a.R) load_structure(row);
    // ---- phase 1: everything the producer wrote
    pdl_wait();
    ktrace_waited(a.trace_slot);
    if (stamp) g_cta_phase[blockIdx.x * 3 + 1] = global_ns();
    // batch-norm vectors of the inputs: warp 0 (self) and warp 1 (cross) derive them - fp64 sums, divide, rsqrt: ~100
    // instructions - and publish them in shared memory; every thread of every warp used to repeat that
    const int warp_id = tid >> 5;
    if (warp_id == 0) ls.issue_acc(a.bn_s);
    if (CROSS && warp_id == 1) lc.issue_acc(a.bn_c);
    float4 xs_raw = f4_zero();
    if (row < a.R) {
        xs_raw = ld4(a.Xs + (size_t)rr * 4);
        if (NCSR > 0) ga.load_rows(a.Xs);
        if (CROSS) gb.load_rows(a.Xc);
    }
    if (warp_id == 0) {
        if (ls.mode == 2 && !(a.ablate & 2)) {          // batch statistics: one feature per lane
            float sc, sh, mu, rs;
            bn4_lane(ls.v, ls.w, ls.b, a.bn_s.inv_n, sc, sh, mu, rs);
            if (tid < 4) { bnv[tid] = sc; bnv[4 + tid] = sh; }
        } else {
            Bn4 v;
            if (a.ablate & 2) { v.sc = make_float4(1.f, 1.f, 1.f, 1.f); v.sh = f4_zero(); }
            else v = ls.resolve(a.bn_s);
            if (tid == 0) { *reinterpret_cast<float4*>(bnv) = v.sc; *reinterpret_cast<float4*>(bnv + 4) = v.sh; }
        }
    } else if (CROSS && warp_id == 1) {
        if (lc.mode == 2 && !(a.ablate & 2)) {
            float sc, sh, mu, rs;
            bn4_lane(lc.v, lc.w, lc.b, a.bn_c.inv_n, sc, sh, mu, rs);
            if ((tid & 31) < 4) { bnv[8 + (tid & 31)] = sc; bnv[12 + (tid & 31)] = sh; }
        } else {
            Bn4 v;
            if (a.ablate & 2) { v.sc = make_float4(1.f, 1.f, 1.f, 1.f); v.sh = f4_zero(); }
            else v = lc.resolve(a.bn_c);
            if ((tid & 31) == 0) { *reinterpret_cast<float4*>(bnv + 8) = v.sc; *reinterpret_cast<float4*>(bnv + 12) = v.sh; }
        }
    }
    __syncthreads();                                   // weights and batch-norm vectors in shared memory
    if (stamp) g_cta_phase[blockIdx.x * 3 + 2] = global_ns();
    const float4 sc_s = *reinterpret_cast<const float4*>(bnv), sh_s = *reinterpret_cast<const float4*>(bnv + 4);
    const float4 sc_c = CROSS ? *reinterpret_cast<const float4*>(bnv + 8) : sc_s;
    const float4 sh_c = CROSS ? *reinterpret_cast<const float4*>(bnv + 12) : sh_s;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};

    for (bool first = true; row < a.R; row += stride, first = false) {
        if (!first) {
            load_structure(row);
            xs_raw = ld4(a.Xs + (size_t)rr * 4);
            if (NCSR > 0) ga.load_rows(a.Xs);
            if (CROSS) gb.load_rows(a.Xc);
        }
        if (rw <= 0.f) continue;                       // a skipped copy of a phantom line-graph row
        float4 x1[NB];
        const float4 xs = f4_affine(xs_raw, sc_s, sh_s);
        x1[0] = xs;
        x1[1] = make_float4(d * xs.x, d * xs.y, d * xs.z, d * xs.w);
        // sum val*(s*z+t) = s*(sum val*z) + t*(sum val)
        float4 acc0 = f4_zero(), am = f4_zero(), ad = f4_zero(), unused4 = f4_zero();
        float ws0 = 0.f, wm = 0.f, wd = 0.f, unused = 0.f;
        if (NCSR > 0) ga.accumulate(acc0, ws0, unused4, unused);
        if (CROSS) gb.accumulate(am, wm, ad, wd);
        if (NCSR > 0)       // long rows: the remaining entries, batch by batch
            for (int k = k0[0] + BA; k < k1[0]; k += BA) {
                GatherBatch<BA, false> g;
                g.load_entries(a.col[0], a.val[0], nullptr, k, k1[0]);
                g.load_rows(a.Xs);
                g.accumulate(acc0, ws0, unused4, unused);
            }
        if (CROSS)
            for (int k = p0 + BP; k < p1; k += BP) {
                GatherBatch<BP, true> g;
                g.load_entries(a.p_col, a.p_pm, a.p_pd, k, p1);
                g.load_rows(a.Xc);
                g.accumulate(am, wm, ad, wd);
            }
        if (NCSR > 0)
            x1[2] = make_float4(fmaf(acc0.x, sc_s.x, ws0 * sh_s.x), fmaf(acc0.y, sc_s.y, ws0 * sh_s.y),
                                fmaf(acc0.z, sc_s.z, ws0 * sh_s.z), fmaf(acc0.w, sc_s.w, ws0 * sh_s.w));
#pragma unroll
        for (int t = 1; t < NCSR; ++t) {
            float4 acc = f4_zero();
            float ws = 0.f;
            csr_gather4(a.col[t], a.val[t], k0[t], k1[t], a.Xs, acc, ws);
            x1[2 + t] = make_float4(fmaf(acc.x, sc_s.x, ws * sh_s.x), fmaf(acc.y, sc_s.y, ws * sh_s.y),
                                    fmaf(acc.z, sc_s.z, ws * sh_s.z), fmaf(acc.w, sc_s.w, ws * sh_s.w));
        }
        if (CROSS) {
            x1[2 + NCSR] = make_float4(fmaf(am.x, sc_c.x, wm * sh_c.x), fmaf(am.y, sc_c.y, wm * sh_c.y),
                                       fmaf(am.z, sc_c.z, wm * sh_c.z), fmaf(am.w, sc_c.w, wm * sh_c.w));
            x1[3 + NCSR] = make_float4(fmaf(ad.x, sc_c.x, wd * sh_c.x), fmaf(ad.y, sc_c.y, wd * sh_c.y),
                                       fmaf(ad.z, sc_c.z, wd * sh_c.z), fmaf(ad.w, sc_c.w, wd * sh_c.w));
        }
        if (a.X1) {
#pragma unroll
            for (int b = 0; b < NB; ++b) *reinterpret_cast<float4*>(a.X1 + ((size_t)rr * NB + b) * 4) = x1[b];
        }
        // mat-vec: 4 outputs x (NB*4) inputs, weights broadcast from shared memory
        float out[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = bias[o];
#pragma unroll
            for (int b = 0; b < NB; ++b)
                acc += f4_dot(x1[b], *reinterpret_cast<const float4*>(W + (o * NB + b) * 4));
            if (o >= a.relu_from) acc = fmaxf(acc, 0.f);
            out[o] = acc;
            s1[o] = fmaf(rw, acc, s1[o]);
            s2[o] = fmaf(rw * acc, acc, s2[o]);
        }
        *reinterpret_cast<float4*>(a.Z + (size_t)rr * 4) = make_float4(out[0], out[1], out[2], out[3]);
    }
    if (stamp) g_cta_phase[blockIdx.x * 3] = global_ns();
    if (a.acc_out && !(a.ablate & 4)) {   // warp shuffle tree -> one row per warp in shared memory -> 8 fp64 atomics per CTA
        const int lane = tid & 31, warp = tid >> 5;
        double st[8];
#pragma unroll
        for (int o = 0; o < 4; ++o) { st[o] = (double)s1[o]; st[4 + o] = (double)s2[o]; }
        warp_reduce_scatter<double, 8>(st);              // lane l: total of value l >> 2
        if ((lane & 3) == 0) red[warp * 8 + (lane >> 2)] = st[0];
        __syncthreads();
        if (tid < 8) {
            double v = 0.0;
            for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * 8 + tid];
            accum_add(a.acc_out, 8, hgnn_ws_bins(8), tid, v);
        }
    }
    if (stamp) g_cta_times[blockIdx.x * 3 + 1] = global_ns();
    ktrace_end(a.trace_slot);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct Bwd4Args {
    // side being differentiated (width 4)
    const float* gY; const float* Z; int relu_from; int Rg; int has_bn;
    const double* acc_f; const double* acc_b; const float* bn_w;
    const float* Wa; int Ha; const float* Wb; int Hb; int Cin;
    double* dW_bins; double* db_bins;
    // self part: transposed [IDENT, DIAG, CSR...]
    int R_self, n_csr;
    const float* diag;
    const int* rowptr[2]; const int* col[2]; const float* val[2];
    const int* rng_rowptr; const int* rng_id; const float* rng_val; const int* rng_lo; const int* rng_hi;  // op 0 only
    const float* Xs; BnRef bn_s; float* gXs; int acc_self; double* acc_b_self;
    // cross part: Pm^T / Pd^T pattern with rows = rows of the cross tensor
    int R_cross;
    const int* pt_rowptr; const int* pt_col; const float* pt_pm; const float* pt_pd;
    const float* Xc; BnRef bn_c; float* gXc; int acc_cross; double* acc_b_cross;
    int col0_cross;
    const float* roww_s; const float* roww_c;   // optional row weights of the self / cross rows (see Fwd4Args::roww)
    const int* rowmap_s; const int* rowmap_c;   // optional row lists (R_self / R_cross entries) of the two parts
    int ctas_self;        // CTAs [0, ctas_self) work on the self rows
    // dedicated range-sum CTAs (optional): the last `range_ctas` CTAs of the grid publish the sums of the rng_n ranges
    // into rng_sum_g (4 floats each) and set rng_flag_g[r]; both zero on entry
    int rng_n, range_ctas; float* rng_sum_g; int* rng_flag_g;
    int ablate;           // timing experiments only (HGNN_B200_ABLATE): 1 no gathers, 2 no range phase, 4 no flush, 8 CTA times
    int trace_slot;       // >= 0: record this launch in g_ktrace (HGNN_B200_ABLATE bit 16)
    double inv_Rg;        // 1 / Rg (host-divided)
};

struct Gpre4 {
    float4 c0, c1, c2;
    int relu_from;
    bool bn, need_z;
    const float* G;
    const float* Z;
    __device__ __forceinline__ float4 operator()(int row) const {
        float4 g = ld4(G + (size_t)row * 4);
        if (!need_z) return g;
        const float4 z = ld4(Z + (size_t)row * 4);
        if (bn) g = make_float4(fmaf(c2.x, z.x, fmaf(c0.x, g.x, c1.x)), fmaf(c2.y, z.y, fmaf(c0.y, g.y, c1.y)),
                                fmaf(c2.z, z.z, fmaf(c0.z, g.z, c1.z)), fmaf(c2.w, z.w, fmaf(c0.w, g.w, c1.w)));
        if (0 >= relu_from && !(z.x > 0.f)) g.x = 0.f;
        if (1 >= relu_from && !(z.y > 0.f)) g.y = 0.f;
        if (2 >= relu_from && !(z.z > 0.f)) g.z = 0.f;
        if (3 >= relu_from && !(z.w > 0.f)) g.w = 0.f;
        return g;
    }
};

// sum_k val[k] * gpre(col[k]); batches of GB entries (two row loads per entry), predicated
template <int GB>
__device__ __forceinline__ float4 gpre_gather(const Gpre4& gp, const int* __restrict__ col,
                                              const float* __restrict__ val, int k0, int k1) {
    float4 acc = f4_zero();
    for (int k = k0; k < k1; k += GB) {
        int c[GB];
        float v[GB];
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const bool on = k + j < k1;
            c[j] = on ? __ldg(col + k + j) : -1;
            v[j] = on ? __ldg(val + k + j) : 0.f;
        }
        float4 g[GB];
#pragma unroll
        for (int j = 0; j < GB; ++j) g[j] = c[j] >= 0 ? gp(c[j]) : f4_zero();
#pragma unroll
        for (int j = 0; j < GB; ++j) acc = f4_fma(v[j], g[j], acc);
    }
    return acc;
}

#define R4_MAX_FLAGGED 64
#define R4_MAX_IDS 8

// DW = false: the gather-only variant (the weight gradients come from dw_row4_kernel, which streams
// over the x1 rows saved by the forward on a parallel graph branch): ~48 fewer live registers.
// profiling aid (HGNN_B200_ABLATE bit 8): per-CTA start / end / role of the last backward launch

// GB / CB: entries per gather batch of the self part (first transposed operator) / of the cross part,
// picked by the host from the average row lengths so that a typical row needs one batch.
#ifndef R4_BWD_MIN_CTAS
#define R4_BWD_MIN_CTAS 4     // <= 128 registers = 4 CTAs per SM; without the bound ptxas takes 153 (3 CTAs per SM: 1.01 instead of 0.96 ms per step)
#endif
template <int NCSR, bool DW, int GB, int CB>
__global__ void __launch_bounds__(R4_THREADS, R4_BWD_MIN_CTAS)
bwd_row4_kernel(const Bwd4Args a) {
    constexpr int NT = 2 + NCSR;                          // self blocks
    __shared__ __align__(16) float Ws[NT * 4 * 4];        // [t][o][f] = W[o][t*4+f]
    __shared__ __align__(16) float Wc[2 * 4 * 4];         // [t][o][f] = W[o][col0 + t*4 + f]
    __shared__ float red[(R4_THREADS / 32) * 64];
    __shared__ int flagged[R4_MAX_FLAGGED];
    __shared__ int n_flagged;
    __shared__ int rng_ids[R4_MAX_IDS];
    __shared__ float rng_sum[R4_MAX_IDS * 4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_self = (int)blockIdx.x < a.ctas_self;
    const int row_ctas = gridDim.x - a.range_ctas;          // CTAs [row_ctas, gridDim.x) only compute range sums
    const bool is_range = (int)blockIdx.x >= row_ctas;
    pdl_launch_dependents();
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) {
        g_cta_times[blockIdx.x * 3] = global_ns();
        g_cta_times[blockIdx.x * 3 + 2] = is_self ? 1 : 0;
    }
    // ---- phase 0 (parameters only; under PDL this overlaps the producer's tail)
    if (tid == 0) n_flagged = 0;
    for (int i = tid; i < NT * 16; i += R4_THREADS) {
        const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Ws[i] = wrow[t * 4 + f];
    }
    if (a.R_cross > 0)
        for (int i = tid; i < 32; i += R4_THREADS) {
            const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
            const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
            Wc[i] = wrow[a.col0_cross + t * 4 + f];
        }
    // Graph structure of the thread's FIRST row (row list, weight, diagonal, row pointers): nothing the producer writes,
    // so these three dependent rounds run before the wait, under the producer's tail.
    int pre_row = -1, pre_k0 = 0, pre_k1 = 0;
    float pre_rw = 1.f, pre_d = 0.f;
    if (!is_range) {
        if (is_self) {
            const int ridx = blockIdx.x * R4_THREADS + tid;
            if (ridx < a.R_self) {
                pre_row = a.rowmap_s ? __ldg(a.rowmap_s + ridx) : ridx;
                pre_rw = a.roww_s ? __ldg(a.roww_s + pre_row) : 1.f;
                pre_d = __ldg(a.diag + pre_row);
                if (NCSR > 0) { pre_k0 = __ldg(a.rowptr[0] + pre_row); pre_k1 = __ldg(a.rowptr[0] + pre_row + 1); }
            }
        } else {
            const int ridx = (blockIdx.x - a.ctas_self) * R4_THREADS + tid;
            if (ridx < a.R_cross) {
                pre_row = a.rowmap_c ? __ldg(a.rowmap_c + ridx) : ridx;
                pre_rw = a.roww_c ? __ldg(a.roww_c + pre_row) : 1.f;
                pre_k0 = __ldg(a.pt_rowptr + pre_row);
                pre_k1 = __ldg(a.pt_rowptr + pre_row + 1);
            }
        }
    }
    pdl_wait();
    // ---- coefficients of this side's BN + ReLU backward (warp 0) and the input's BN vectors (warp 1), once per CTA,
    //      published in shared memory (every thread used to derive both: ~300 instructions with the fp64 sums)
    __shared__ __align__(16) float cv[28];             // c0, c1, c2 | scale, shift, mean, 1/std
    Gpre4 gp;
    gp.relu_from = a.relu_from; gp.bn = a.has_bn != 0; gp.need_z = a.has_bn != 0 || a.relu_from < 4;
    gp.G = a.gY; gp.Z = a.Z;
    if (warp == 0) {
        float c0 = 1.f, c1 = 0.f, c2 = 0.f;
        if (a.has_bn) {
            double tf[8], tb[8];
            warp_totals8(a.acc_f, tf);
            warp_totals8(a.acc_b, tb);
            const float w = a.bn_w[0];
            const double inv_n = a.inv_Rg;
            const int f = lane & 3;
            const double m = tf[f] * inv_n;
            const double var = fma(-m, m, tf[4 + f] * inv_n);
            const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
            const float k0 = w * r_;
            const float k2 = -k0 * (float)(tb[4 + f] * inv_n) * r_;
            c0 = k0;
            c2 = k2;
            c1 = -k0 * (float)(tb[f] * inv_n) - k2 * (float)m;
        }
        if (lane < 4) { cv[lane] = c0; cv[4 + lane] = c1; cv[8 + lane] = c2; }
    } else if (warp == 1) {
        const Bn4 bx = bn4_from_ref(is_self ? a.bn_s : a.bn_c);
        if (lane == 0) {
            *reinterpret_cast<float4*>(cv + 12) = bx.sc; *reinterpret_cast<float4*>(cv + 16) = bx.sh;
            *reinterpret_cast<float4*>(cv + 20) = bx.mu; *reinterpret_cast<float4*>(cv + 24) = bx.rs;
        }
    }
    __syncthreads();                                   // weights and coefficient vectors in shared memory
    gp.c0 = *reinterpret_cast<const float4*>(cv); gp.c1 = *reinterpret_cast<const float4*>(cv + 4);
    gp.c2 = *reinterpret_cast<const float4*>(cv + 8);
    const float4 sc = *reinterpret_cast<const float4*>(cv + 12), sh = *reinterpret_cast<const float4*>(cv + 16);
    const float4 mu = *reinterpret_cast<const float4*>(cv + 20), rs = *reinterpret_cast<const float4*>(cv + 24);
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_phase[blockIdx.x * 3 + 2] = global_ns();   // coefficients ready

    if (is_range) {
        // ---- a range CTA: sum gPre over each of its ranges (4 rows = 8 loads in flight per thread), publish the
        //      sum, then the flag.  The row CTAs read them after their row loops, microseconds later.
        for (int r = blockIdx.x - row_ctas; r < a.rng_n; r += a.range_ctas) {
            const int lo = __ldg(a.rng_lo + r), hi = __ldg(a.rng_hi + r);
            float4 acc = f4_zero();
            for (int rr = lo + tid; rr < hi; rr += 4 * R4_THREADS) {
                float4 gv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) gv[u] = rr + u * R4_THREADS < hi ? gp(rr + u * R4_THREADS) : f4_zero();
#pragma unroll
                for (int u = 0; u < 4; ++u) { acc.x += gv[u].x; acc.y += gv[u].y; acc.z += gv[u].z; acc.w += gv[u].w; }
            }
            acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
            if (lane == 0) { red[warp * 4] = acc.x; red[warp * 4 + 1] = acc.y; red[warp * 4 + 2] = acc.z; red[warp * 4 + 3] = acc.w; }
            __syncthreads();
            if (tid < 4) {
                float v = 0.f;
                for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * 4 + tid];
                a.rng_sum_g[(size_t)r * 4 + tid] = v;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(a.rng_flag_g + r, 1);
        }
        if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_times[blockIdx.x * 3 + 1] = global_ns();
        return;
    }

    // per-thread accumulators: dW (NT or 2 blocks of 4x4), dbias (4), (sum g, sum g*xhat) (8)
    float dw[NT * 16];
#pragma unroll
    for (int i = 0; i < NT * 16; ++i) dw[i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};

    if (is_self) {
        float* const gX = a.gXs;
        const bool stats = a.acc_b_self != nullptr && gX != nullptr;
        // finishes a row that owns range entries: the delta of everything that is linear in T[2]
        auto range_fixup = [&](int row, const float (&dT)[4]) {
            const float4 xr = ld4(a.Xs + (size_t)row * 4);
            const float4 xn = f4_affine(xr, sc, sh);
            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + (2 * 4 + o) * 4);
                g[0] = fmaf(dT[o], w.x, g[0]); g[1] = fmaf(dT[o], w.y, g[1]);
                g[2] = fmaf(dT[o], w.z, g[2]); g[3] = fmaf(dT[o], w.w, g[3]);
                if (DW) {
                    dw[(2 * 4 + o) * 4 + 0] = fmaf(dT[o], xn.x, dw[(2 * 4 + o) * 4 + 0]);
                    dw[(2 * 4 + o) * 4 + 1] = fmaf(dT[o], xn.y, dw[(2 * 4 + o) * 4 + 1]);
                    dw[(2 * 4 + o) * 4 + 2] = fmaf(dT[o], xn.z, dw[(2 * 4 + o) * 4 + 2]);
                    dw[(2 * 4 + o) * 4 + 3] = fmaf(dT[o], xn.w, dw[(2 * 4 + o) * 4 + 3]);
                }
            }
            if (gX) {
                float4 old = __ldcg(reinterpret_cast<const float4*>(gX + (size_t)row * 4));
                old.x += g[0]; old.y += g[1]; old.z += g[2]; old.w += g[3];
                *reinterpret_cast<float4*>(gX + (size_t)row * 4) = old;
                if (stats) {
                    const float xh[4] = {(xr.x - mu.x) * rs.x, (xr.y - mu.y) * rs.y, (xr.z - mu.z) * rs.z, (xr.w - mu.w) * rs.w};
#pragma unroll
                    for (int f = 0; f < 4; ++f) { sg[f] += g[f]; sgx[f] = fmaf(g[f], xh[f], sgx[f]); }
                }
            }
        };
        const int ridx0 = blockIdx.x * R4_THREADS + tid;
        for (int ridx = ridx0; ridx < a.R_self; ridx += a.ctas_self * R4_THREADS) {
            const bool pre = ridx == ridx0;             // structure of the first row was loaded before the wait
            const int row = pre ? pre_row : (a.rowmap_s ? __ldg(a.rowmap_s + ridx) : ridx);
            const float rw = pre ? pre_rw : (a.roww_s ? __ldg(a.roww_s + row) : 1.f);
            if (rw <= 0.f) continue;                    // a skipped copy of a phantom line-graph row
            float4 T[NT];
            T[0] = gp(row);
            const float d = pre ? pre_d : __ldg(a.diag + row);
            T[1] = make_float4(d * T[0].x, d * T[0].y, d * T[0].z, d * T[0].w);
#pragma unroll
            for (int t = 0; t < NCSR; ++t) {
                int k0, k1;
                if (t == 0 && pre) { k0 = pre_k0; k1 = pre_k1; }
                else { k0 = __ldg(a.rowptr[t] + row); k1 = __ldg(a.rowptr[t] + row + 1); }
                if (a.ablate & 1) k1 = k0;
                T[2 + t] = t == 0 ? gpre_gather<GB>(gp, a.col[t], a.val[t], k0, k1)
                                  : gpre_gather<2>(gp, a.col[t], a.val[t], k0, k1);
            }
            if (NCSR > 0 && a.rng_rowptr && !(a.ablate & 2) && __ldg(a.rng_rowptr + row + 1) > __ldg(a.rng_rowptr + row)) {
                const int slot = atomicAdd(&n_flagged, 1);
                if (slot < R4_MAX_FLAGGED) {
                    flagged[slot] = row;
                } else {            // list full: add the run-length part serially (correct, slow, rare)
                    for (int e = __ldg(a.rng_rowptr + row); e < __ldg(a.rng_rowptr + row + 1); ++e) {
                        const int id = __ldg(a.rng_id + e);
                        const float v = __ldg(a.rng_val + e);
                        for (int rr = __ldg(a.rng_lo + id); rr < __ldg(a.rng_hi + id); ++rr) T[2] = f4_fma(v, gp(rr), T[2]);
                    }
                }
            }
            const float4 xr = ld4(a.Xs + (size_t)row * 4);
            const float4 xn1 = f4_affine(xr, sc, sh);
            const float4 xn = make_float4(rw * xn1.x, rw * xn1.y, rw * xn1.z, rw * xn1.w);   // weight of the row in dW
            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float4 w = *reinterpret_cast<const float4*>(Ws + (t * 4 + o) * 4);
                    g[0] = fmaf(Tv[o], w.x, g[0]); g[1] = fmaf(Tv[o], w.y, g[1]);
                    g[2] = fmaf(Tv[o], w.z, g[2]); g[3] = fmaf(Tv[o], w.w, g[3]);
                    if (DW) {
                        dw[(t * 4 + o) * 4 + 0] = fmaf(Tv[o], xn.x, dw[(t * 4 + o) * 4 + 0]);
                        dw[(t * 4 + o) * 4 + 1] = fmaf(Tv[o], xn.y, dw[(t * 4 + o) * 4 + 1]);
                        dw[(t * 4 + o) * 4 + 2] = fmaf(Tv[o], xn.z, dw[(t * 4 + o) * 4 + 2]);
                        dw[(t * 4 + o) * 4 + 3] = fmaf(Tv[o], xn.w, dw[(t * 4 + o) * 4 + 3]);
                    }
                }
            }
            if (DW) { db[0] = fmaf(rw, T[0].x, db[0]); db[1] = fmaf(rw, T[0].y, db[1]); db[2] = fmaf(rw, T[0].z, db[2]); db[3] = fmaf(rw, T[0].w, db[3]); }
            if (gX) {
                float4 o4 = make_float4(g[0], g[1], g[2], g[3]);
                if (a.acc_self) {
                    const float4 old = *reinterpret_cast<const float4*>(gX + (size_t)row * 4);
                    o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w;
                }
                *reinterpret_cast<float4*>(gX + (size_t)row * 4) = o4;
                if (stats) {
                    const float xh[4] = {(xr.x - mu.x) * rs.x, (xr.y - mu.y) * rs.y, (xr.z - mu.z) * rs.z, (xr.w - mu.w) * rs.w};
#pragma unroll
                    for (int f = 0; f < 4; ++f) { sg[f] = fmaf(rw, g[f], sg[f]); sgx[f] = fmaf(rw * g[f], xh[f], sgx[f]); }
                }
            }
        }
        if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_phase[blockIdx.x * 3] = global_ns();
        // ---- run-length parts of the flagged rows.  (1) the distinct range ids they refer to (a CTA's
        //      rows lie in one or two graphs, so one or two ids), (2) each range sum once per CTA,
        //      cooperatively, (3) every flagged row finished by its own thread, in parallel: the delta of
        //      everything that is linear in T[2].
        __syncthreads();
        const int nf = min(n_flagged, R4_MAX_FLAGGED);
        if (nf > 0 && a.rng_flag_g) {
            // the sums come from the range CTAs of this launch: wait (bounded) for the flag, else sum the range here
            if (tid < nf) {
                const int row = flagged[tid];
                float dT[4] = {0.f, 0.f, 0.f, 0.f};
                for (int e = __ldg(a.rng_rowptr + row); e < __ldg(a.rng_rowptr + row + 1); ++e) {
                    const int id = __ldg(a.rng_id + e);
                    const float v = __ldg(a.rng_val + e);
                    volatile int* flag = a.rng_flag_g + id;
                    int spins = 0;
                    while (*flag == 0 && spins < 100000) ++spins;
                    if (*flag != 0) {
                        __threadfence();
                        const float4 sum = __ldcg(reinterpret_cast<const float4*>(a.rng_sum_g + (size_t)id * 4));
                        dT[0] = fmaf(v, sum.x, dT[0]); dT[1] = fmaf(v, sum.y, dT[1]);
                        dT[2] = fmaf(v, sum.z, dT[2]); dT[3] = fmaf(v, sum.w, dT[3]);
                    } else {        // never seen in practice: the range CTAs are resident from the start of the launch
                        for (int rr = __ldg(a.rng_lo + id); rr < __ldg(a.rng_hi + id); ++rr) {
                            const float4 gv = gp(rr);
                            dT[0] = fmaf(v, gv.x, dT[0]); dT[1] = fmaf(v, gv.y, dT[1]);
                            dT[2] = fmaf(v, gv.z, dT[2]); dT[3] = fmaf(v, gv.w, dT[3]);
                        }
                    }
                }
                range_fixup(row, dT);
            }
        } else if (nf > 0) {
            if (tid < R4_MAX_IDS) rng_ids[tid] = -1;
            __syncthreads();
            int my_row = -1, e0 = 0, e1 = 0;
            if (tid < nf) {
                my_row = flagged[tid];
                e0 = __ldg(a.rng_rowptr + my_row);
                e1 = __ldg(a.rng_rowptr + my_row + 1);
                for (int e = e0; e < e1; ++e) {
                    const int id = __ldg(a.rng_id + e);
                    for (int j = 0; j < R4_MAX_IDS; ++j) {
                        const int old = atomicCAS(&rng_ids[j], -1, id);
                        if (old == -1 || old == id) break;
                    }
                }
            }
            __syncthreads();
            for (int j = 0; j < R4_MAX_IDS; ++j) {
                const int id = rng_ids[j];
                if (id < 0) break;                     // uniform
                const int lo = __ldg(a.rng_lo + id), hi = __ldg(a.rng_hi + id);
                float4 acc = f4_zero();
                for (int rr = lo + tid; rr < hi; rr += 4 * R4_THREADS) {      // 4 rows (8 loads) in flight per thread
                    float4 gv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) gv[u] = rr + u * R4_THREADS < hi ? gp(rr + u * R4_THREADS) : f4_zero();
#pragma unroll
                    for (int u = 0; u < 4; ++u) { acc.x += gv[u].x; acc.y += gv[u].y; acc.z += gv[u].z; acc.w += gv[u].w; }
                }
                acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
                if (lane == 0) { red[warp * 4] = acc.x; red[warp * 4 + 1] = acc.y; red[warp * 4 + 2] = acc.z; red[warp * 4 + 3] = acc.w; }
                __syncthreads();
                if (tid < 4) {
                    float v = 0.f;
                    for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * 4 + tid];
                    rng_sum[j * 4 + tid] = v;
                }
                __syncthreads();
            }
            if (tid < nf) {
                const int row = my_row;
                float dT[4] = {0.f, 0.f, 0.f, 0.f};
                for (int e = e0; e < e1; ++e) {
                    const int id = __ldg(a.rng_id + e);
                    const float v = __ldg(a.rng_val + e);
                    int j = 0;
                    while (j < R4_MAX_IDS && rng_ids[j] != id) ++j;
                    if (j < R4_MAX_IDS) {
#pragma unroll
                        for (int f = 0; f < 4; ++f) dT[f] = fmaf(v, rng_sum[j * 4 + f], dT[f]);
                    } else {        // more distinct ranges than table slots: serial sum (correct, slow, rare)
                        for (int rr = __ldg(a.rng_lo + id); rr < __ldg(a.rng_hi + id); ++rr) {
                            const float4 gv = gp(rr);
                            dT[0] = fmaf(v, gv.x, dT[0]); dT[1] = fmaf(v, gv.y, dT[1]);
                            dT[2] = fmaf(v, gv.z, dT[2]); dT[3] = fmaf(v, gv.w, dT[3]);
                        }
                    }
                }
                range_fixup(row, dT);
            }
        }
        if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) {
            g_cta_phase[blockIdx.x * 3 + 1] = global_ns();
        }
    } else {
        float* const gX = a.gXc;
        const bool stats = a.acc_b_cross != nullptr && gX != nullptr;
        const int ncta = row_ctas - a.ctas_self;
        const int ridx0 = (blockIdx.x - a.ctas_self) * R4_THREADS + tid;
        for (int ridx = ridx0; ridx < a.R_cross; ridx += ncta * R4_THREADS) {
            const bool pre = ridx == ridx0;
            const int row = pre ? pre_row : (a.rowmap_c ? __ldg(a.rowmap_c + ridx) : ridx);
            const float rw = pre ? pre_rw : (a.roww_c ? __ldg(a.roww_c + row) : 1.f);
            if (rw <= 0.f) continue;
            float4 Tm = f4_zero(), Td = f4_zero();
            int k0 = pre ? pre_k0 : __ldg(a.pt_rowptr + row), k1 = pre ? pre_k1 : __ldg(a.pt_rowptr + row + 1);
            if (a.ablate & 1) k1 = k0;
            for (int k = k0; k < k1; k += CB) {         // CB entries (2 CB row loads) in flight
                int c[CB];
                float vm[CB], vd[CB];
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    const bool on = k + j < k1;
                    c[j] = on ? __ldg(a.pt_col + k + j) : -1;
                    vm[j] = on ? __ldg(a.pt_pm + k + j) : 0.f;
                    vd[j] = on ? __ldg(a.pt_pd + k + j) : 0.f;
                }
                float4 gv[CB];
#pragma unroll
                for (int j = 0; j < CB; ++j) gv[j] = c[j] >= 0 ? gp(c[j]) : f4_zero();
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    Tm = f4_fma(vm[j], gv[j], Tm);
                    Td = f4_fma(vd[j], gv[j], Td);
                }
            }
            const float4 xr = ld4(a.Xc + (size_t)row * 4);
            const float4 xn1 = f4_affine(xr, sc, sh);
            const float4 xn = make_float4(rw * xn1.x, rw * xn1.y, rw * xn1.z, rw * xn1.w);
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            const float4 T[2] = {Tm, Td};
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float4 w = *reinterpret_cast<const float4*>(Wc + (t * 4 + o) * 4);
                    g[0] = fmaf(Tv[o], w.x, g[0]); g[1] = fmaf(Tv[o], w.y, g[1]);
                    g[2] = fmaf(Tv[o], w.z, g[2]); g[3] = fmaf(Tv[o], w.w, g[3]);
                    if (DW) {
                        dw[(t * 4 + o) * 4 + 0] = fmaf(Tv[o], xn.x, dw[(t * 4 + o) * 4 + 0]);
                        dw[(t * 4 + o) * 4 + 1] = fmaf(Tv[o], xn.y, dw[(t * 4 + o) * 4 + 1]);
                        dw[(t * 4 + o) * 4 + 2] = fmaf(Tv[o], xn.z, dw[(t * 4 + o) * 4 + 2]);
                        dw[(t * 4 + o) * 4 + 3] = fmaf(Tv[o], xn.w, dw[(t * 4 + o) * 4 + 3]);
                    }
                }
            }
            if (gX) {
                float4 o4 = make_float4(g[0], g[1], g[2], g[3]);
                if (a.acc_cross) {
                    const float4 old = *reinterpret_cast<const float4*>(gX + (size_t)row * 4);
                    o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w;
                }
                *reinterpret_cast<float4*>(gX + (size_t)row * 4) = o4;
                if (stats) {
                    const float xh[4] = {(xr.x - mu.x) * rs.x, (xr.y - mu.y) * rs.y, (xr.z - mu.z) * rs.z, (xr.w - mu.w) * rs.w};
#pragma unroll
                    for (int f = 0; f < 4; ++f) { sg[f] = fmaf(rw, g[f], sg[f]); sgx[f] = fmaf(rw * g[f], xh[f], sgx[f]); }
                }
            }
        }
        if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_phase[blockIdx.x * 3] = g_cta_phase[blockIdx.x * 3 + 1] = global_ns();
    }
    // ---- flush: warp shuffle tree -> per-warp rows in shared memory -> one fp64 atomic per value
    const int nvals = is_self ? NT * 16 : 32;
    const int nbw = hgnn_ws_bins(4 * a.Cin);
    const int col_base = is_self ? 0 : a.col0_cross;
    __syncthreads();
    if (a.ablate & 4) return;
    if (DW) {
        float pad[64];                                   // NT * 16 <= 64 values, zero padded
#pragma unroll
        for (int i = 0; i < 64; ++i) pad[i] = i < NT * 16 ? dw[i] : 0.f;
        warp_reduce_scatter<float, 64>(pad);             // lane l: totals of values 2l, 2l + 1
        red[warp * 64 + 2 * lane] = pad[0];
        red[warp * 64 + 2 * lane + 1] = pad[1];
    }
    __syncthreads();
    if (DW && a.dW_bins)
        for (int i = tid; i < nvals; i += R4_THREADS) {
            float v = 0.f;
            for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * 64 + i];
            const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
            accum_add(a.dW_bins, 4 * a.Cin, nbw, o * a.Cin + col_base + t * 4 + f, (double)v);
        }
    __syncthreads();
    // dbias (self only) and the BN sums of the produced gradient
    float extra[16];
#pragma unroll
    for (int f = 0; f < 4; ++f) { extra[f] = db[f]; extra[4 + f] = sg[f]; extra[8 + f] = sgx[f]; extra[12 + f] = 0.f; }
    warp_reduce_scatter<float, 16>(extra);               // lane l: total of value l >> 1
    if ((lane & 1) == 0) red[warp * 64 + (lane >> 1)] = extra[0];
    __syncthreads();
    if (tid < 12) {
        double v = 0.0;
        for (int w = 0; w < R4_THREADS / 32; ++w) v += (double)red[w * 64 + tid];
        if (tid < 4) {
            if (DW && is_self && a.db_bins) accum_add(a.db_bins, 4, hgnn_ws_bins(4), tid, v);
        } else {
            double* accb = is_self ? a.acc_b_self : a.acc_b_cross;
            float* gXp = is_self ? a.gXs : a.gXc;
            if (accb && gXp) accum_add(accb, 8, hgnn_ws_bins(8), tid - 4, v);
        }
    }
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_times[blockIdx.x * 3 + 1] = global_ns();
}

// ---------------------------------------------------------------------------------------------
// backward with the weight-gradient reduction on the tensor cores (one CSR operator, no run-length part: A^T, or
// the collapsed AL^T)
// ---------------------------------------------------------------------------------------------
// bwd_row4_kernel keeps 48 + 12 per-thread accumulators (dW, dbias, batch-norm sums) alive across its rows and folds
// them at the end with reduce-scatter butterflies: ncu attributes 22 % of all executed warp instructions of the
// kernel to that flush (profiles/README.md round 2, item 10) - paid per warp, and a warp walks only 1-2 rows per
// thread.  The sums are an outer-product reduction over the rows of a warp,
//   dW[t][o][f]   = sum_rows T[t][o] * (w xn[f])        (12 x 4)
//   sum g xhat[f] = sum_rows g[f] * (w xhat[f])         (diagonal of 4 x 4)
// i.e. C (16 x 8) = A^T B with A = [T | g] (32 rows x 16) and B = [w xn | w xhat] (32 rows x 8): exactly one
// mma.m16n8k8 tile, K = the 32 lanes.  So per iteration every lane writes its 24 values into a warp-private
// shared-memory tile (conflict-free: value j of lane l at j * 36 + l), the warp loads the A / B fragments back
// (4 k-steps x 6 LDS), splits them (3xTF32: x = hi + lo, lo*hi + hi*lo + hi*hi - fp32-grade, as in engine_wide.cuh) and
// issues 12 mma.sync; C lives in 4 registers.  ~120 instructions per iteration against 60 FMA + a 450-instruction
// flush; no accumulator arrays (80 registers -> 6 CTAs of 128 threads per SM).  dbias and sum g stay in 8 registers.
#define R4C_STRIDE 36                 // floats per tile column block (32 rows + pad, 16-byte aligned)
#define R4C_VALS 24
template <int GB, int CB>
__global__ void __launch_bounds__(R4_THREADS, 6)
bwd_row4c_kernel(const Bwd4Args a) {
    __shared__ __align__(16) float Ws[3 * 16];               // [t][o][f] = W[o][t*4+f]
    __shared__ __align__(16) float Wc[2 * 16];               // [t][o][f] = W[o][col0 + t*4 + f]
    __shared__ __align__(16) float vec[32];                  // gPre c0,c1,c2 (12) | pad (4) | input sc,sh,mu,rs (16)
    __shared__ __align__(16) float tile[(R4_THREADS / 32) * R4C_VALS * R4C_STRIDE];
    __shared__ float red[(R4_THREADS / 32) * 64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_self = (int)blockIdx.x < a.ctas_self;
    pdl_launch_dependents();
    // ---- parameters only (overlaps the producer's tail under PDL)
    for (int i = tid; i < 48; i += R4_THREADS) {
        const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Ws[i] = wrow[t * 4 + f];
    }
    if (a.R_cross > 0)
        for (int i = tid; i < 32; i += R4_THREADS) {
            const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
            const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
            Wc[i] = wrow[a.col0_cross + t * 4 + f];
        }
    pdl_wait();
    // ---- coefficient vectors into shared memory: warp 0 the BN + ReLU backward of this side, warp 1 the input's BN
    if (warp == 0) {
        float c0 = 1.f, c1 = 0.f, c2 = 0.f;
        if (a.has_bn) {
            double tf[8], tb[8];
            warp_totals8(a.acc_f, tf);
            warp_totals8(a.acc_b, tb);
            const int f = lane & 3;
            const float w = a.bn_w[0];
            const double inv_n = a.inv_Rg;
            const double m = tf[f] * inv_n;
            const double var = fma(-m, m, tf[4 + f] * inv_n);
            const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
            const float k0 = w * r_;
            const float k2 = -k0 * (float)(tb[4 + f] * inv_n) * r_;
            c0 = k0; c2 = k2;
            c1 = -k0 * (float)(tb[f] * inv_n) - k2 * (float)m;
        }
        if (lane < 4) { vec[lane] = c0; vec[4 + lane] = c1; vec[8 + lane] = c2; }
    } else if (warp == 1) {
        const Bn4 bx = bn4_from_ref(is_self ? a.bn_s : a.bn_c);
        if (lane == 0) {
            *reinterpret_cast<float4*>(vec + 16) = bx.sc; *reinterpret_cast<float4*>(vec + 20) = bx.sh;
            *reinterpret_cast<float4*>(vec + 24) = bx.mu; *reinterpret_cast<float4*>(vec + 28) = bx.rs;
        }
    }
    __syncthreads();
    Gpre4 gp;
    gp.relu_from = a.relu_from; gp.bn = a.has_bn != 0; gp.need_z = a.has_bn != 0 || a.relu_from < 4;
    gp.G = a.gY; gp.Z = a.Z;
    gp.c0 = *reinterpret_cast<const float4*>(vec); gp.c1 = *reinterpret_cast<const float4*>(vec + 4);
    gp.c2 = *reinterpret_cast<const float4*>(vec + 8);

    float* const my = tile + warp * (R4C_VALS * R4C_STRIDE);
    const int fg = lane >> 2, ft = lane & 3;            // fragment coordinates of mma.m16n8k8
    float cfr[4] = {0.f, 0.f, 0.f, 0.f};                 // C fragment: (fg, 2ft) (fg, 2ft+1) (fg+8, 2ft) (fg+8, 2ft+1)
    float db[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f};

    const int R = is_self ? a.R_self : a.R_cross;
    const int first = (is_self ? blockIdx.x : blockIdx.x - a.ctas_self) * R4_THREADS;
    const int step = (is_self ? a.ctas_self : (int)gridDim.x - a.ctas_self) * R4_THREADS;
    const int* const rowmap = is_self ? a.rowmap_s : a.rowmap_c;
    const float* const roww = is_self ? a.roww_s : a.roww_c;
    const float* const X = is_self ? a.Xs : a.Xc;
    float* const gX = is_self ? a.gXs : a.gXc;
    const bool accumulate = (is_self ? a.acc_self : a.acc_cross) != 0;
    const bool stats = (is_self ? a.acc_b_self : a.acc_b_cross) != nullptr && gX != nullptr;
    // whole warps iterate together (the tile is filled by all 32 lanes, inactive ones with zeros)
    for (int base = first + warp * 32; base < R; base += step) {
        const int ridx = base + lane;
        float4 T[3] = {f4_zero(), f4_zero(), f4_zero()};
        float4 gq = f4_zero(), xw = f4_zero(), xhw = f4_zero();
        int row = -1;
        float rw = 0.f;
        if (ridx < R) {
            row = rowmap ? __ldg(rowmap + ridx) : ridx;
            rw = roww ? __ldg(roww + row) : 1.f;
            if (rw <= 0.f) row = -1;
        }
        if (row >= 0) {
            const float4 xr = ld4(X + (size_t)row * 4);
            float4 old = f4_zero();
            if (gX && accumulate) old = __ldcg(reinterpret_cast<const float4*>(gX + (size_t)row * 4));
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (is_self) {
                T[0] = gp(row);
                const float d = __ldg(a.diag + row);
                const int k0 = __ldg(a.rowptr[0] + row), k1 = __ldg(a.rowptr[0] + row + 1);
                T[2] = gpre_gather<GB>(gp, a.col[0], a.val[0], k0, k1);
                T[1] = make_float4(d * T[0].x, d * T[0].y, d * T[0].z, d * T[0].w);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const float4 w = *reinterpret_cast<const float4*>(Ws + (t * 4 + o) * 4);
                        g[0] = fmaf(Tv[o], w.x, g[0]); g[1] = fmaf(Tv[o], w.y, g[1]);
                        g[2] = fmaf(Tv[o], w.z, g[2]); g[3] = fmaf(Tv[o], w.w, g[3]);
                    }
                }
                db[0] = fmaf(rw, T[0].x, db[0]); db[1] = fmaf(rw, T[0].y, db[1]);
                db[2] = fmaf(rw, T[0].z, db[2]); db[3] = fmaf(rw, T[0].w, db[3]);
            } else {
                const int k0 = __ldg(a.pt_rowptr + row), k1 = __ldg(a.pt_rowptr + row + 1);
                for (int k = k0; k < k1; k += CB) {         // CB entries (2 CB row loads) in flight
                    int c[CB];
                    float vm[CB], vd[CB];
#pragma unroll
                    for (int j = 0; j < CB; ++j) {
                        const bool on = k + j < k1;
                        c[j] = on ? __ldg(a.pt_col + k + j) : -1;
                        vm[j] = on ? __ldg(a.pt_pm + k + j) : 0.f;
                        vd[j] = on ? __ldg(a.pt_pd + k + j) : 0.f;
                    }
                    float4 gv[CB];
#pragma unroll
                    for (int j = 0; j < CB; ++j) gv[j] = c[j] >= 0 ? gp(c[j]) : f4_zero();
#pragma unroll
                    for (int j = 0; j < CB; ++j) {
                        T[0] = f4_fma(vm[j], gv[j], T[0]);
                        T[1] = f4_fma(vd[j], gv[j], T[1]);
                    }
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const float4 w = *reinterpret_cast<const float4*>(Wc + (t * 4 + o) * 4);
                        g[0] = fmaf(Tv[o], w.x, g[0]); g[1] = fmaf(Tv[o], w.y, g[1]);
                        g[2] = fmaf(Tv[o], w.z, g[2]); g[3] = fmaf(Tv[o], w.w, g[3]);
                    }
                }
            }
            const float4 sc = *reinterpret_cast<const float4*>(vec + 16), sh = *reinterpret_cast<const float4*>(vec + 20);
            const float4 xn = f4_affine(xr, sc, sh);
            xw = make_float4(rw * xn.x, rw * xn.y, rw * xn.z, rw * xn.w);
            if (gX) {
                *reinterpret_cast<float4*>(gX + (size_t)row * 4) =
                    make_float4(g[0] + old.x, g[1] + old.y, g[2] + old.z, g[3] + old.w);
                if (stats) {
                    const float4 mu = *reinterpret_cast<const float4*>(vec + 24), rs = *reinterpret_cast<const float4*>(vec + 28);
                    gq = make_float4(g[0], g[1], g[2], g[3]);
                    xhw = make_float4(rw * (xr.x - mu.x) * rs.x, rw * (xr.y - mu.y) * rs.y,
                                      rw * (xr.z - mu.z) * rs.z, rw * (xr.w - mu.w) * rs.w);
                    sg[0] = fmaf(rw, g[0], sg[0]); sg[1] = fmaf(rw, g[1], sg[1]);
                    sg[2] = fmaf(rw, g[2], sg[2]); sg[3] = fmaf(rw, g[3], sg[3]);
                }
            }
        }
        // ---- per-row values -> warp tile (value j of lane l at my[j * STRIDE + l]), then two products per lane
        __syncwarp();
        {
            const float v[R4C_VALS] = {T[0].x, T[0].y, T[0].z, T[0].w, T[1].x, T[1].y, T[1].z, T[1].w,
                                       T[2].x, T[2].y, T[2].z, T[2].w, gq.x, gq.y, gq.z, gq.w,
                                       xw.x, xw.y, xw.z, xw.w, xhw.x, xhw.y, xhw.z, xhw.w};
#pragma unroll
            for (int j = 0; j < R4C_VALS; ++j) my[j * R4C_STRIDE + lane] = v[j];
        }
        __syncwarp();
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {                 // two pairs of k-steps (8 lanes = 8 rows each)
            uint32_t ah[2][4], al[2][4];
            SplitB bb[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = 8 * (2 * pr + h) + ft;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    tf32_split(my[(fg + 8 * (e & 1)) * R4C_STRIDE + k0 + 4 * (e >> 1)], ah[h][e], al[h][e]);
                bb[h] = split_b(my[(16 + fg) * R4C_STRIDE + k0], my[(16 + fg) * R4C_STRIDE + k0 + 4]);
            }
            mma_3xtf32_fresh_pair(cfr, ah[0], al[0], bb[0], ah[1], al[1], bb[1]);
        }
    }
    // ---- flush: the C fragment holds the warp totals; dbias / sum g by shuffles.  red[warp]: [0, 48) dW (m * 4 + f),
    //      [48, 52) sum g xhat, [52, 56) dbias, [56, 60) sum g
    float extra[8];
#pragma unroll
    for (int f = 0; f < 4; ++f) { extra[f] = db[f]; extra[4 + f] = sg[f]; }
    warp_reduce_scatter<float, 8>(extra);                // lane l: total of value l >> 2
    if (ft < 2) {                                        // columns 0..3 = dW
        red[warp * 64 + fg * 4 + 2 * ft] = cfr[0];
        red[warp * 64 + fg * 4 + 2 * ft + 1] = cfr[1];
        if (fg < 4) {
            red[warp * 64 + (fg + 8) * 4 + 2 * ft] = cfr[2];
            red[warp * 64 + (fg + 8) * 4 + 2 * ft + 1] = cfr[3];
        }
    } else if (fg >= 4) {                                // rows 12..15 x columns 4..7: the diagonal is sum g xhat
        const int f = fg - 4;                            // row 12 + f = fg + 8, column 4 + f
        if (2 * ft == 4 + f) red[warp * 64 + 48 + f] = cfr[2];
        if (2 * ft + 1 == 4 + f) red[warp * 64 + 48 + f] = cfr[3];
    }
    if ((lane & 3) == 0) red[warp * 64 + 52 + (lane >> 2)] = extra[0];
    __syncthreads();
    const int nbw = hgnn_ws_bins(4 * a.Cin);
    const int col_base = is_self ? 0 : a.col0_cross;
    if (tid < 60) {
        double v = 0.0;
        for (int w = 0; w < R4_THREADS / 32; ++w) v += (double)red[w * 64 + tid];
        if (tid < 48) {
            const int t = tid >> 4, o = (tid >> 2) & 3, f = tid & 3;
            if ((is_self || t < 2) && a.dW_bins) accum_add(a.dW_bins, 4 * a.Cin, nbw, o * a.Cin + col_base + t * 4 + f, v);
        } else if (tid < 52) {              // sum g xhat
            double* accb = is_self ? a.acc_b_self : a.acc_b_cross;
            if (stats && accb) accum_add(accb, 8, hgnn_ws_bins(8), 4 + (tid - 48), v);
        } else if (tid < 56) {              // dbias (self part only)
            if (is_self && a.db_bins) accum_add(a.db_bins, 4, hgnn_ws_bins(4), tid - 52, v);
        } else {                            // sum g
            double* accb = is_self ? a.acc_b_self : a.acc_b_cross;
            if (stats && accb) accum_add(accb, 8, hgnn_ws_bins(8), tid - 56, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradients as a streaming pass:  dW[o][c] = sum_rows gPre[row][o] * x1[row][c],
// dbias[o] = sum_rows gPre[row][o], with x1 saved by the forward.  No gathers, no dependent loads;
// runs on a parallel graph branch, off the critical path of the backward chain.
// ---------------------------------------------------------------------------------------------
struct Dw4Args {
    const float* gY; const float* Z; int relu_from; int R; int has_bn;
    const double* acc_f; const double* acc_b; const float* bn_w;
    const float* X1; int Cin;
    double* dW_bins; double* db_bins;
};

template <int NB>
__global__ void __launch_bounds__(R4_THREADS)
dw_row4_kernel(const Dw4Args a) {
    __shared__ float red[(R4_THREADS / 32) * (16 * NB + 4)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Gpre4 gp;
    gp.relu_from = a.relu_from; gp.bn = a.has_bn != 0; gp.need_z = a.has_bn != 0 || a.relu_from < 4;
    gp.G = a.gY; gp.Z = a.Z;
    if (a.has_bn) {
        double tf[8], tb[8];
        warp_totals8(a.acc_f, tf);
        warp_totals8(a.acc_b, tb);
        const float w = a.bn_w[0];
        const double inv_n = 1.0 / (double)a.R;
        float c0[4], c1[4], c2[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const double m = tf[f] * inv_n;
            const double var = fma(-m, m, tf[4 + f] * inv_n);
            const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
            const float k0 = w * r_;
            const float k2 = -k0 * (float)(tb[4 + f] * inv_n) * r_;
            c0[f] = k0; c2[f] = k2;
            c1[f] = -k0 * (float)(tb[f] * inv_n) - k2 * (float)m;
        }
        gp.c0 = make_float4(c0[0], c0[1], c0[2], c0[3]);
        gp.c1 = make_float4(c1[0], c1[1], c1[2], c1[3]);
        gp.c2 = make_float4(c2[0], c2[1], c2[2], c2[3]);
    } else {
        gp.c0 = make_float4(1.f, 1.f, 1.f, 1.f); gp.c1 = f4_zero(); gp.c2 = f4_zero();
    }
    float dw[16 * NB];
#pragma unroll
    for (int i = 0; i < 16 * NB; ++i) dw[i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    for (int row = blockIdx.x * R4_THREADS + tid; row < a.R; row += gridDim.x * R4_THREADS) {
        const float4 g = gp(row);
        float4 x1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) x1[b] = ld4(a.X1 + ((size_t)row * NB + b) * 4);
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            db[o] += gv[o];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                dw[(o * NB + b) * 4 + 0] = fmaf(gv[o], x1[b].x, dw[(o * NB + b) * 4 + 0]);
                dw[(o * NB + b) * 4 + 1] = fmaf(gv[o], x1[b].y, dw[(o * NB + b) * 4 + 1]);
                dw[(o * NB + b) * 4 + 2] = fmaf(gv[o], x1[b].z, dw[(o * NB + b) * 4 + 2]);
                dw[(o * NB + b) * 4 + 3] = fmaf(gv[o], x1[b].w, dw[(o * NB + b) * 4 + 3]);
            }
        }
    }
    constexpr int NV = 16 * NB + 4;
#pragma unroll
    for (int i = 0; i < 16 * NB; ++i) {
        const float v = warp_sum(dw[i]);
        if (lane == 0) red[warp * NV + i] = v;
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const float v = warp_sum(db[o]);
        if (lane == 0) red[warp * NV + 16 * NB + o] = v;
    }
    __syncthreads();
    const int nbw = hgnn_ws_bins(4 * a.Cin);
    for (int i = tid; i < NV; i += R4_THREADS) {
        double v = 0.0;
        for (int w = 0; w < R4_THREADS / 32; ++w) v += (double)red[w * NV + i];
        if (i < 16 * NB) accum_add(a.dW_bins, 4 * a.Cin, nbw, i, v);      // i = o*Cin + c  (Cin = 4*NB)
        else if (a.db_bins) accum_add(a.db_bins, 4, hgnn_ws_bins(4), i - 16 * NB, v);
    }
}
