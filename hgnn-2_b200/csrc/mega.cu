// mega.cu -- persistent cooperative kernels: a whole run of width-4 layer sides of GNN_simple / GNN_lg per launch.
//
// Replaces, for the script-default state width (h = 2: every state row is one float4), the chain of per-side
// launches of engine_row4.cuh (2 x 18 sides per pass at L = 20) by ONE kernel per pass.  Reference semantics:
// models/layers/layers_mnb.py:52-69 (layer_simple), :189-225 / :256-290 / :322-358 (layer_with_lg_1/2/3) with
// the batch-norm of models/layers/batch_normalization.py:34-43,65-93, exactly as the per-side kernels compute
// them (raw activations, consumers normalise on load from the producer's binned fp64 sums).
//
// Why one kernel: at this width a side moves <= 25 MB, all of it L2-resident, and a per-side launch spends most
// of its 7-18 us in things that are not the gather - launch ramp / drain, parameter and row-pointer round
// trips, a cold L1 for the graph structure (profiles/README.md, ablation table).  Here
//   * every CTA owns a FIXED slice of the node rows and of the active line-graph rows for all layers, so the
//     row pointers / columns / values of its slice are re-read from its own L1 (ld.global.nc; L1 survives for
//     the whole launch), only activations (ld.global.cg) come from L2;
//   * sides are separated by a grid barrier (~2 us measured, profiles/logs/r2a_gb_probe.log) instead of a
//     launch boundary; the batch-norm sums ride on it (fp64 atomics before, one 32-double load after);
//   * phantom line-graph rows are computed once per graph (multiplicity weights, sparse_ops._build_collapsed):
//     half of the line-graph rows and 85 % of nnz(AL) disappear, and with them the run-length machinery.
// HBM-/L2-latency-bound fp32 + integer work: no tensor cores (the linear is 20 x 4 per row).
#include "common.cuh"
#include "mega.cuh"

namespace mk {

constexpr float BN_EPS = 1e-5f;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 acc) {
    acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
    return acc;
}
__device__ __forceinline__ float4 f4_affine(float4 x, float4 s, float4 t) {
    return make_float4(fmaf(x.x, s.x, t.x), fmaf(x.y, s.y, t.y), fmaf(x.z, s.z, t.z), fmaf(x.w, s.w, t.w));
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
// sum val*(s*z+t) over a row's entries = s*(sum val*z) + t*(sum val)
__device__ __forceinline__ float4 f4_affine_sum(float4 acc, float ws, float4 s, float4 t) {
    return make_float4(fmaf(acc.x, s.x, ws * t.x), fmaf(acc.y, s.y, ws * t.y), fmaf(acc.z, s.z, ws * t.z), fmaf(acc.w, s.w, ws * t.w));
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// profiling aid: time stamp k (0: phase entered, 1: barrier passed, 2: vectors ready, 3: rows + flush done)
__device__ __forceinline__ void trace_mark(const Params& P, int phase, int k) {
    if (P.trace && threadIdx.x == 0) P.trace[((size_t)phase * gridDim.x + blockIdx.x) * 4 + k] = global_ns();
}

__device__ __forceinline__ int slice_lo(int n, int b, int G) { return (int)(((long long)n * b) / G); }

// ---- grid barrier: ONE monotonic arrival counter.  Thread 0 of every CTA arrives with a fire-and-forget
// red.release (orders the CTA's earlier writes and atomics - bar.sync makes them cumulative) and polls the
// counter with ld.acquire until all CTAs of the round are in.  The counter is 0 at launch: the CTA that
// finishes the kernel last resets it (grid_exit).
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        while ((int)(ld_acquire(bar) - target) < 0) {}
    }
    __syncthreads();
}
__device__ __forceinline__ void grid_exit(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev;
        asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar + 32) : "memory");
        if (prev == gridDim.x - 1) {         // everybody is past its last poll
            asm volatile("st.relaxed.gpu.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
            asm volatile("st.relaxed.gpu.u32 [%0], %1;" ::"l"(bar + 32), "r"(0u) : "memory");
        }
    }
}

// exclusive prefix sum of v over the CTA (thread order); total = sum over all threads.  scratch: THREADS/32 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* scratch, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = lane < THREADS / 32 ? scratch[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL, t, o);
            if (lane >= o) t += u;
        }
        if (lane < THREADS / 32) scratch[lane] = t;
    }
    __syncthreads();
    const int base = warp > 0 ? scratch[warp - 1] : 0;
    total = scratch[THREADS / 32 - 1];
    return base + incl - v;
}

// ---- thread-private structure cache in shared memory ---------------------------------------------------
// A grid barrier ends in an acquire at gpu scope, which invalidates the SM's L1 (CCTL.IVALL): row pointers,
// columns and values re-read through L1 after every barrier come from L2 again - four dependent round trips
// per row and phase (profiles/logs/mega_trace_r2d.log: 4.4 us of row work per forward phase, 15-26 us per
// backward phase).  Every thread owns the SAME rows in every phase, so at kernel start it copies the structure
// of its rows into shared memory - header (row, weight, diagonal, counts, offset) + the entries of its two
// operator sets - and from then on a row costs one shared-memory lookup plus ONE round of feature loads from L2.
// Rows that do not fit the budget keep offset -1 and read their structure from global memory.
struct StageSrc {
    const float* diag;
    const int* rp1; const int* col1; const float* val1;                      // CSR operator of this row space
    const int* rp2; const int* col2; const float* pm2; const float* pd2;      // incidence rows (NULL: none)
};
template <int KIND>
__device__ __forceinline__ StageSrc stage_src(const Graph& g, bool bwd) {
    StageSrc s;
    if (KIND == 0) {
        s.diag = g.deg;
        s.rp1 = bwd ? g.at_rp : g.a_rp; s.col1 = bwd ? g.at_col : g.a_col; s.val1 = bwd ? g.at_val : g.a_val;
        s.rp2 = g.p_rp; s.col2 = g.p_col; s.pm2 = g.p_pm; s.pd2 = g.p_pd;
    } else {
        s.diag = g.dl;
        s.rp1 = bwd ? g.btc_rp : g.b_rp; s.col1 = bwd ? g.btc_col : g.b_col; s.val1 = bwd ? g.btc_val : g.b_val;
        s.rp2 = g.pt_rp; s.col2 = g.pt_col; s.pm2 = g.pt_pm; s.pd2 = g.pt_pd;
    }
    return s;
}
constexpr int HDR_N = 4;    // node row header:  d, off, n1, n2
constexpr int HDR_E = 6;    // line-graph row header: row, w, d, off, n1, n2

template <int KIND>
__device__ __forceinline__ void stage_rows(const Graph& g, bool bwd, int* hdr, int* ent, int cap, int& used, int* scratch) {
    const StageSrc src = stage_src<KIND>(g, bwd);
    const int n = KIND == 0 ? g.Rn : g.n_act;
    const int lo = slice_lo(n, blockIdx.x, gridDim.x), hi = slice_lo(n, blockIdx.x + 1, gridDim.x);
    int words = 0;
    for (int i = lo + (int)threadIdx.x; i < hi; i += THREADS) {
        const int row = KIND == 0 ? i : __ldg(g.erow + i);
        const int n1 = __ldg(src.rp1 + row + 1) - __ldg(src.rp1 + row);
        const int n2 = src.rp2 ? __ldg(src.rp2 + row + 1) - __ldg(src.rp2 + row) : 0;
        words += 2 * n1 + 3 * n2;
    }
    int total = 0;
    int base = used + block_excl_scan(words, scratch, total);
    for (int i = lo + (int)threadIdx.x; i < hi; i += THREADS) {
        const int row = KIND == 0 ? i : __ldg(g.erow + i);
        const int k1 = __ldg(src.rp1 + row), n1 = __ldg(src.rp1 + row + 1) - k1;
        const int k2 = src.rp2 ? __ldg(src.rp2 + row) : 0, n2 = src.rp2 ? __ldg(src.rp2 + row + 1) - k2 : 0;
        const int w_row = 2 * n1 + 3 * n2;
        const int off = base + w_row <= cap ? base : -1;
        base += w_row;
        int* h = hdr + (size_t)(i - lo) * (KIND == 0 ? HDR_N : HDR_E);
        if (KIND == 0) {
            h[0] = __float_as_int(__ldg(src.diag + row)); h[1] = off; h[2] = n1; h[3] = n2;
        } else {
            h[0] = row; h[1] = __float_as_int(__ldg(g.ew + row)); h[2] = __float_as_int(__ldg(src.diag + row));
            h[3] = off; h[4] = n1; h[5] = n2;
        }
        if (off >= 0) {
            int* e = ent + off;
            for (int k = 0; k < n1; ++k) { e[k] = __ldg(src.col1 + k1 + k); e[n1 + k] = __float_as_int(__ldg(src.val1 + k1 + k)); }
            e += 2 * n1;
            for (int k = 0; k < n2; ++k) {
                e[k] = __ldg(src.col2 + k2 + k);
                e[n2 + k] = __float_as_int(__ldg(src.pm2 + k2 + k));
                e[2 * n2 + k] = __float_as_int(__ldg(src.pd2 + k2 + k));
            }
        }
    }
    used += total;
}

struct RowView {
    int row, n1, n2;
    float w, d;
    const int* c1; const float* v1;
    const int* c2; const float* m2; const float* d2;
};
template <int KIND>
__device__ __forceinline__ RowView row_view(const Graph& g, bool bwd, const int* hdr, const int* ent, int li, int i) {
    RowView r;
    const int* h = hdr + (size_t)li * (KIND == 0 ? HDR_N : HDR_E);
    int off;
    if (KIND == 0) {
        r.row = i; r.w = 1.f; r.d = __int_as_float(h[0]); off = h[1]; r.n1 = h[2]; r.n2 = h[3];
    } else {
        r.row = h[0]; r.w = __int_as_float(h[1]); r.d = __int_as_float(h[2]); off = h[3]; r.n1 = h[4]; r.n2 = h[5];
    }
    if (off >= 0) {
        const int* e = ent + off;
        r.c1 = e; r.v1 = reinterpret_cast<const float*>(e + r.n1);
        e += 2 * r.n1;
        r.c2 = e; r.m2 = reinterpret_cast<const float*>(e + r.n2); r.d2 = reinterpret_cast<const float*>(e + 2 * r.n2);
    } else {
        const StageSrc src = stage_src<KIND>(g, bwd);
        const int k1 = __ldg(src.rp1 + r.row);
        r.c1 = src.col1 + k1; r.v1 = src.val1 + k1;
        const int k2 = src.rp2 ? __ldg(src.rp2 + r.row) : 0;
        r.c2 = src.col2 + k2; r.m2 = src.pm2 + k2; r.d2 = src.pd2 + k2;
    }
    return r;
}

// Sum of NV per-lane values over the warp with a reduce-scatter butterfly (~NV shuffles instead of 5 NV).
// Afterwards lane l holds the totals of  NV = 64: 2l, 2l+1 in val[0], val[1];  32: l;  16: l >> 1;  8: l >> 2.
template <typename T, int NV>
__device__ __forceinline__ void warp_reduce_scatter(T (&val)[NV]) {
    const int lane = threadIdx.x & 31;
    int n = NV;
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) {
        const int mask = 16 >> bit;
        const bool upper = (lane & mask) != 0;
        if (n > 1) {
            const int half = n >> 1;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < half) {
                    const T send = upper ? val[i] : val[i + half];
                    const T keep = upper ? val[i + half] : val[i];
                    val[i] = keep + __shfl_xor_sync(FULL, send, mask);
                }
            }
            n = half;
        } else {
            val[0] += __shfl_xor_sync(FULL, val[0], mask);
        }
    }
}

// ---- batch-norm vectors of a width-4 tensor, computed by ONE warp into shared memory -------------------
// out[0..3] scale, [4..7] shift, [8..11] mean, [12..15] 1/std.  Same arithmetic as engine_row4.cuh
// (E[x^2] - mean^2 in fp64, the rest in fp32).
__device__ __forceinline__ void bn_vectors(const Tensor& t, float* out) {
    const int lane = threadIdx.x & 31;
    if (!t.acc_f) {
        if (lane < 4) { out[lane] = 1.f; out[4 + lane] = 0.f; out[8 + lane] = 0.f; out[12 + lane] = 1.f; }
        return;
    }
    double v = hgnn_bins8_lane(t.acc_f);           // hgnn_ws_bins(8) x 8 doubles
    v += __shfl_xor_sync(FULL, v, 8);
    v += __shfl_xor_sync(FULL, v, 16);
    const int f = lane & 3;
    const double sum = __shfl_sync(FULL, v, f), sq = __shfl_sync(FULL, v, 4 + f);
    if (lane < 4) {
        const float w = __ldg(t.bn_w), b = __ldg(t.bn_b);
        const double inv_n = t.inv_n;
        const double m = sum * inv_n;
        const double var = fma(-m, m, sq * inv_n);
        const float r = 1.0f / sqrtf(fmaxf((float)var, 0.f) + BN_EPS);
        out[lane] = w * r;
        out[4 + lane] = b - w * (float)m * r;
        out[8 + lane] = (float)m;
        out[12 + lane] = r;
    }
}

// coefficients of  gPre = (c0 g + c1 + c2 z) * relu_mask  for the tensor being differentiated (one warp)
__device__ __forceinline__ void gpre_vectors(const Tensor& t, float* out) {
    const int lane = threadIdx.x & 31;
    if (!t.acc_b) {
        if (lane < 4) { out[lane] = 1.f; out[4 + lane] = 0.f; out[8 + lane] = 0.f; }
        return;
    }
    double vf = hgnn_bins8_lane(t.acc_f), vb = hgnn_bins8_lane(t.acc_b);
    vf += __shfl_xor_sync(FULL, vf, 8); vf += __shfl_xor_sync(FULL, vf, 16);
    vb += __shfl_xor_sync(FULL, vb, 8); vb += __shfl_xor_sync(FULL, vb, 16);
    const int f = lane & 3;
    const double sum = __shfl_sync(FULL, vf, f), sq = __shfl_sync(FULL, vf, 4 + f);
    const double sg = __shfl_sync(FULL, vb, f), sgx = __shfl_sync(FULL, vb, 4 + f);
    if (lane < 4) {
        const float w = __ldg(t.bn_w);
        const double inv_n = t.inv_n;
        const double m = sum * inv_n;
        const double var = fma(-m, m, sq * inv_n);
        const float r = 1.0f / sqrtf(fmaxf((float)var, 0.f) + BN_EPS);
        const float k0 = w * r;
        const float k2 = -k0 * (float)(sgx * inv_n) * r;
        out[lane] = k0;
        out[4 + lane] = -k0 * (float)(sg * inv_n) - k2 * (float)m;
        out[8 + lane] = k2;
    }
}

// One batch of a CSR gather: indices / values (structure: L1), then the feature rows (activations: L2).
template <int B, bool TWO>
struct GatherBatch {
    int c[B];
    float v[B], v2[TWO ? B : 1];
    float4 x[B];
    // col / val point into the thread's shared-memory cache (or, for unstaged rows, into global memory)
    __device__ __forceinline__ void load_entries(const int* col, const float* val, const float* val2, int k, int k1) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const bool on = k + j < k1;
            c[j] = on ? col[k + j] : -1;
            v[j] = on ? val[k + j] : 0.f;
            if (TWO) v2[j] = on ? val2[k + j] : 0.f;
        }
    }
    __device__ __forceinline__ void load_rows(const float* X) {
#pragma unroll
        for (int j = 0; j < B; ++j) x[j] = c[j] >= 0 ? ldcg4(X + (size_t)c[j] * 4) : f4_zero();
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

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
// KIND 0: node rows (operators I, D = deg, A; cross = Pm / Pd over the line-graph tensor)
// KIND 1: active line-graph rows (I, D = dl, AL; cross = Pm^T / Pd^T over the node tensor), weighted by ew
template <int KIND, bool CROSS, int BA, int BP>
__device__ __forceinline__ void fwd_side(const Params& P, const Side& sd, const float* W, const float* bias,
                                         const float* bnv, const int* hdr, const int* ent,
                                         float (&s1)[4], float (&s2)[4]) {
    constexpr int NB = 3 + (CROSS ? 2 : 0);
    const Graph& g = P.g;
    const float* Xs = P.t[sd.src_self].data;
    const float* Xc = CROSS ? P.t[sd.src_cross].data : nullptr;
    float* Z = P.t[sd.out].data;
    const int n = KIND == 0 ? g.Rn : g.n_act;
    const int lo = slice_lo(n, blockIdx.x, gridDim.x), hi = slice_lo(n, blockIdx.x + 1, gridDim.x);
    const float4 sc_s = lds4(bnv), sh_s = lds4(bnv + 4), sc_c = lds4(bnv + 16), sh_c = lds4(bnv + 20);
    for (int i = lo + (int)threadIdx.x; i < hi; i += THREADS) {
        const RowView r = row_view<KIND>(g, false, hdr, ent, i - lo, i);
        GatherBatch<BA, false> ga;
        GatherBatch<BP, true> gb;
        ga.load_entries(r.c1, r.v1, nullptr, 0, r.n1);
        if (CROSS) gb.load_entries(r.c2, r.m2, r.d2, 0, r.n2);
        const float4 xs_raw = ldcg4(Xs + (size_t)r.row * 4);
        ga.load_rows(Xs);
        if (CROSS) gb.load_rows(Xc);
        float4 x1[NB];
        const float4 xs = f4_affine(xs_raw, sc_s, sh_s);
        x1[0] = xs;
        x1[1] = make_float4(r.d * xs.x, r.d * xs.y, r.d * xs.z, r.d * xs.w);
        float4 acc0 = f4_zero(), am = f4_zero(), ad = f4_zero(), u4 = f4_zero();
        float ws0 = 0.f, wm = 0.f, wd = 0.f, u = 0.f;
        ga.accumulate(acc0, ws0, u4, u);
        if (CROSS) gb.accumulate(am, wm, ad, wd);
        for (int k = BA; k < r.n1; k += BA) {          // long rows: the remaining entries, batch by batch
            GatherBatch<BA, false> t;
            t.load_entries(r.c1, r.v1, nullptr, k, r.n1);
            t.load_rows(Xs);
            t.accumulate(acc0, ws0, u4, u);
        }
        if (CROSS)
            for (int k = BP; k < r.n2; k += BP) {
                GatherBatch<BP, true> t;
                t.load_entries(r.c2, r.m2, r.d2, k, r.n2);
                t.load_rows(Xc);
                t.accumulate(am, wm, ad, wd);
            }
        x1[2] = f4_affine_sum(acc0, ws0, sc_s, sh_s);
        if (CROSS) {
            x1[3] = f4_affine_sum(am, wm, sc_c, sh_c);
            x1[4] = f4_affine_sum(ad, wd, sc_c, sh_c);
        }
        float out[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = bias[o];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc += f4_dot(x1[b], lds4(W + (o * NB + b) * 4));
            if (o >= sd.relu_from) acc = fmaxf(acc, 0.f);
            out[o] = acc;
            s1[o] = fmaf(r.w, acc, s1[o]);
            s2[o] = fmaf(r.w * acc, acc, s2[o]);
        }
        *reinterpret_cast<float4*>(Z + (size_t)r.row * 4) = make_float4(out[0], out[1], out[2], out[3]);
    }
}

// skipped line-graph rows of tensor `buf` <- their representative (ew[r] = -(distance to it))
__device__ __forceinline__ void expand_rows(const Graph& g, float* buf) {
    const int lo = slice_lo(g.Rm, blockIdx.x, gridDim.x), hi = slice_lo(g.Rm, blockIdx.x + 1, gridDim.x);
    for (int r = lo + (int)threadIdx.x; r < hi; r += THREADS) {
        const float e = __ldg(g.ew + r);
        if (e < 0.f) *reinterpret_cast<float4*>(buf + (size_t)r * 4) = ldcg4(buf + (size_t)(r + (int)e) * 4);
    }
}

#define MK_WSTRIDE 84      // per side: W[4][Cin <= 20] + bias[4]

// dynamic shared memory: [node row headers | line-graph row headers | entries]
extern __shared__ __align__(16) int mk_dyn[];

__device__ __forceinline__ void stage_all(const Params& P, bool bwd, int*& hdrN, int*& hdrE, int*& ent, int* scratch) {
    hdrN = mk_dyn;
    hdrE = hdrN + (size_t)P.max_n * HDR_N;
    ent = hdrE + (size_t)P.max_e * HDR_E;
    int used = 0;
    stage_rows<0>(P.g, bwd, hdrN, ent, P.cap_words, used, scratch);
    if (P.g.n_act > 0) stage_rows<1>(P.g, bwd, hdrE, ent, P.cap_words, used, scratch);
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1) mega_fwd_kernel(const __grid_constant__ Params P) {
    __shared__ __align__(16) float Wsm[MAX_SIDES * MK_WSTRIDE];
    __shared__ __align__(16) float bnv[32];              // [self | cross] x (scale, shift, mean, 1/std)
    __shared__ double red[(THREADS / 32) * 8];
    __shared__ int scratch[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < P.n_sides * MK_WSTRIDE; i += THREADS) {
        const int s = i / MK_WSTRIDE, j = i - s * MK_WSTRIDE;
        const Side& sd = P.s[s];
        float v = 0.f;
        if (j < 4 * sd.Cin) {
            const int o = j / sd.Cin, c = j - o * sd.Cin;
            v = o < sd.Ha ? sd.Wa[(size_t)o * sd.Cin + c] : sd.Wb[(size_t)(o - sd.Ha) * sd.Cin + c];
        } else if (j >= 80) {
            const int o = j - 80;
            v = o < sd.Ha ? (sd.ba ? sd.ba[o] : 0.f) : (sd.bb ? sd.bb[o - sd.Ha] : 0.f);
        }
        Wsm[i] = v;
    }
    int *hdrN, *hdrE, *ent;
    stage_all(P, false, hdrN, hdrE, ent, scratch);
    unsigned target = 0;
    for (int s = 0; s < P.n_sides; ++s) {
        const Side& sd = P.s[s];
        trace_mark(P, s, 0);
        if (s > 0) grid_sync(P.bar, target);
        trace_mark(P, s, 1);
        const bool cross = sd.src_cross >= 0;
        if (warp == 0) bn_vectors(P.t[sd.src_self], bnv);
        else if (warp == 1 && cross) bn_vectors(P.t[sd.src_cross], bnv + 16);
        __syncthreads();
        trace_mark(P, s, 2);
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        const float* W = Wsm + s * MK_WSTRIDE;
        if (sd.kind == 0) {
            if (cross) fwd_side<0, true, 4, 8>(P, sd, W, W + 80, bnv, hdrN, ent, s1, s2);
            else fwd_side<0, false, 8, 1>(P, sd, W, W + 80, bnv, hdrN, ent, s1, s2);
        } else {
            if (cross) fwd_side<1, true, 4, 4>(P, sd, W, W + 80, bnv, hdrE, ent, s1, s2);
            else fwd_side<1, false, 4, 1>(P, sd, W, W + 80, bnv, hdrE, ent, s1, s2);
        }
        // (sum w z, sum w z^2) of this CTA -> 8 fp64 atomics into the producer bins of the output tensor
        double st[8];
#pragma unroll
        for (int o = 0; o < 4; ++o) { st[o] = (double)s1[o]; st[4 + o] = (double)s2[o]; }
        warp_reduce_scatter<double, 8>(st);              // lane l: total of value l >> 2
        if ((lane & 3) == 0) red[warp * 8 + (lane >> 2)] = st[0];
        __syncthreads();
        if (tid < 8) {
            double v = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) v += red[w * 8 + tid];
            atomicAdd(const_cast<double*>(P.t[sd.out].acc_f) + (size_t)(blockIdx.x & 3) * 8 + tid, v);
        }
        trace_mark(P, s, 3);
    }
    if (P.expand >= 0) {
        grid_sync(P.bar, target);
        expand_rows(P.g, P.t[P.expand].data);
    }
    grid_exit(P.bar);
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
struct Gpre {
    float4 c0, c1, c2;
    int relu_from;
    bool bn, need_z;
    const float* G;
    const float* Z;
    __device__ __forceinline__ float4 operator()(int row) const {
        float4 g = ldcg4(G + (size_t)row * 4);
        if (!need_z) return g;
        const float4 z = ldcg4(Z + (size_t)row * 4);
        if (bn) g = make_float4(fmaf(c2.x, z.x, fmaf(c0.x, g.x, c1.x)), fmaf(c2.y, z.y, fmaf(c0.y, g.y, c1.y)),
                                fmaf(c2.z, z.z, fmaf(c0.z, g.z, c1.z)), fmaf(c2.w, z.w, fmaf(c0.w, g.w, c1.w)));
        if (0 >= relu_from && !(z.x > 0.f)) g.x = 0.f;
        if (1 >= relu_from && !(z.y > 0.f)) g.y = 0.f;
        if (2 >= relu_from && !(z.z > 0.f)) g.z = 0.f;
        if (3 >= relu_from && !(z.w > 0.f)) g.w = 0.f;
        return g;
    }
};

// sum_k val[k] * gpre(col[k]) (and the second value array on the same pattern); batches of GB entries
template <int GB, bool TWO>
__device__ __forceinline__ void gpre_gather(const Gpre& gp, const int* col, const float* val, const float* val2,
                                            int k0, int k1, float4& a1, float4& a2) {
    for (int k = k0; k < k1; k += GB) {
        int c[GB];
        float v[GB], v2[TWO ? GB : 1];
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const bool on = k + j < k1;
            c[j] = on ? col[k + j] : -1;
            v[j] = on ? val[k + j] : 0.f;
            if (TWO) v2[j] = on ? val2[k + j] : 0.f;
        }
        float4 g[GB];
#pragma unroll
        for (int j = 0; j < GB; ++j) g[j] = c[j] >= 0 ? gp(c[j]) : f4_zero();
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            a1 = f4_fma(v[j], g[j], a1);
            if (TWO) a2 = f4_fma(v2[j], g[j], a2);
        }
    }
}

// per-thread partial sums of one part -> the binned fp64 accumulators of the step arena
// dw: NT blocks of 4 x 4 ([t][o][f]); db: sum gPre (self part only); sg / sgx: batch-norm sums of the produced gradient
template <int NT>
__device__ __forceinline__ void flush_part(float (&dw)[NT * 16], const float (&db)[4], const float (&sg)[4],
                                           const float (&sgx)[4], float* red, const Side& sd, int col_base,
                                           bool with_db, double* accb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NV = NT * 16 <= 32 ? 32 : 64;
    float pad[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pad[i] = i < NT * 16 ? dw[i] : 0.f;
    warp_reduce_scatter<float, NV>(pad);
    __syncthreads();                                      // red may still be read by the previous flush
    if (NV == 64) { red[warp * 80 + 2 * lane] = pad[0]; red[warp * 80 + 2 * lane + 1] = pad[1]; }
    else red[warp * 80 + lane] = pad[0];
    float extra[16];
#pragma unroll
    for (int f = 0; f < 4; ++f) { extra[f] = db[f]; extra[4 + f] = sg[f]; extra[8 + f] = sgx[f]; extra[12 + f] = 0.f; }
    warp_reduce_scatter<float, 16>(extra);                // lane l: total of value l >> 1
    if ((lane & 1) == 0) red[warp * 80 + 64 + (lane >> 1)] = extra[0];
    __syncthreads();
    const int nbw = hgnn_ws_bins(4 * sd.Cin);
    if (tid < NT * 16) {
        float v = 0.f;
        for (int w = 0; w < THREADS / 32; ++w) v += red[w * 80 + tid];
        const int t = tid >> 4, o = (tid >> 2) & 3, f = tid & 3;
        atomicAdd(sd.dW_bins + (size_t)(blockIdx.x & (nbw - 1)) * (4 * sd.Cin) + o * sd.Cin + col_base + t * 4 + f, (double)v);
    } else if (tid >= 64 && tid < 76) {
        const int i = tid - 64;
        double v = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) v += (double)red[w * 80 + 64 + i];
        if (i < 4) {
            if (with_db && sd.db_bins) atomicAdd(sd.db_bins + (size_t)(blockIdx.x & (hgnn_ws_bins(4) - 1)) * 4 + i, v);
        } else if (accb) {
            atomicAdd(accb + (size_t)(blockIdx.x & 3) * 8 + (i - 4), v);
        }
    }
}

// Self part of side `sd`: rows of its self input; transposed operators [I, D, CSR^T] = entry set 1 of the row.
template <int KIND, int GB>
__device__ __forceinline__ void bwd_self(const Params& P, const Side& sd, const Gpre& gp, const float* Ws,
                                         const float* bi, const int* hdr, const int* ent, float* red) {
    const Graph& g = P.g;
    const Tensor& tx = P.t[sd.src_self];
    const float* X = tx.data;
    float* gX = sd.need_self ? tx.grad : nullptr;
    double* accb = gX ? tx.acc_b : nullptr;
    const int n = KIND == 0 ? g.Rn : g.n_act;
    const int lo = slice_lo(n, blockIdx.x, gridDim.x), hi = slice_lo(n, blockIdx.x + 1, gridDim.x);
    const float4 sc = lds4(bi), sh = lds4(bi + 4), mu = lds4(bi + 8), rs = lds4(bi + 12);
    float dw[48];
#pragma unroll
    for (int i = 0; i < 48; ++i) dw[i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lo + (int)threadIdx.x; i < hi; i += THREADS) {
        const RowView r = row_view<KIND>(g, true, hdr, ent, i - lo, i);
        const int row = r.row;
        const float w = r.w;
        float4 T[3];
        T[0] = gp(row);
        const float4 xr = ldcg4(X + (size_t)row * 4);
        float4 old = f4_zero();
        if (gX && sd.acc_self) old = ldcg4(gX + (size_t)row * 4);
        T[2] = f4_zero();
        float4 unused = f4_zero();
        gpre_gather<GB, false>(gp, r.c1, r.v1, nullptr, 0, r.n1, T[2], unused);
        T[1] = make_float4(r.d * T[0].x, r.d * T[0].y, r.d * T[0].z, r.d * T[0].w);
        const float4 xn = f4_affine(xr, sc, sh);
        const float xw[4] = {w * xn.x, w * xn.y, w * xn.z, w * xn.w};
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float4 wv = lds4(Ws + (t * 4 + o) * 4);
                gv[0] = fmaf(Tv[o], wv.x, gv[0]); gv[1] = fmaf(Tv[o], wv.y, gv[1]);
                gv[2] = fmaf(Tv[o], wv.z, gv[2]); gv[3] = fmaf(Tv[o], wv.w, gv[3]);
#pragma unroll
                for (int f = 0; f < 4; ++f) dw[(t * 4 + o) * 4 + f] = fmaf(Tv[o], xw[f], dw[(t * 4 + o) * 4 + f]);
            }
        }
        db[0] = fmaf(w, T[0].x, db[0]); db[1] = fmaf(w, T[0].y, db[1]);
        db[2] = fmaf(w, T[0].z, db[2]); db[3] = fmaf(w, T[0].w, db[3]);
        if (gX) {
            *reinterpret_cast<float4*>(gX + (size_t)row * 4) =
                make_float4(gv[0] + old.x, gv[1] + old.y, gv[2] + old.z, gv[3] + old.w);
            if (accb) {
                const float xh[4] = {(xr.x - mu.x) * rs.x, (xr.y - mu.y) * rs.y, (xr.z - mu.z) * rs.z, (xr.w - mu.w) * rs.w};
#pragma unroll
                for (int f = 0; f < 4; ++f) { sg[f] = fmaf(w, gv[f], sg[f]); sgx[f] = fmaf(w * gv[f], xh[f], sgx[f]); }
            }
        }
    }
    flush_part<3>(dw, db, sg, sgx, red, sd, 0, true, accb);
}

// Cross part of side `sd`: rows of its cross input (the OTHER row space), incidence entries (set 2) of those rows.
template <int KIND, int CB>
__device__ __forceinline__ void bwd_cross(const Params& P, const Side& sd, const Gpre& gp, const float* Wc,
                                          const float* bi, const int* hdr, const int* ent, float* red) {
    const Graph& g = P.g;
    const Tensor& tx = P.t[sd.src_cross];
    const float* X = tx.data;
    float* gX = sd.need_cross ? tx.grad : nullptr;
    double* accb = gX ? tx.acc_b : nullptr;
    // node side: cross tensor lives on the line graph (rows = active line-graph rows, pattern Pm^T / Pd^T);
    // edge side: cross tensor lives on the nodes (rows = nodes, pattern Pm / Pd)
    constexpr int RK = KIND == 0 ? 1 : 0;
    const int n = RK == 1 ? g.n_act : g.Rn;
    const int lo = slice_lo(n, blockIdx.x, gridDim.x), hi = slice_lo(n, blockIdx.x + 1, gridDim.x);
    const float4 sc = lds4(bi), sh = lds4(bi + 4), mu = lds4(bi + 8), rs = lds4(bi + 12);
    float dw[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) dw[i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lo + (int)threadIdx.x; i < hi; i += THREADS) {
        const RowView r = row_view<RK>(g, true, hdr, ent, i - lo, i);
        const int row = r.row;
        const float w = r.w;
        const float4 xr = ldcg4(X + (size_t)row * 4);
        float4 old = f4_zero();
        if (gX && sd.acc_cross) old = ldcg4(gX + (size_t)row * 4);
        float4 T[2] = {f4_zero(), f4_zero()};
        gpre_gather<CB, true>(gp, r.c2, r.m2, r.d2, 0, r.n2, T[0], T[1]);
        const float4 xn = f4_affine(xr, sc, sh);
        const float xw[4] = {w * xn.x, w * xn.y, w * xn.z, w * xn.w};
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float4 wv = lds4(Wc + (t * 4 + o) * 4);
                gv[0] = fmaf(Tv[o], wv.x, gv[0]); gv[1] = fmaf(Tv[o], wv.y, gv[1]);
                gv[2] = fmaf(Tv[o], wv.z, gv[2]); gv[3] = fmaf(Tv[o], wv.w, gv[3]);
#pragma unroll
                for (int f = 0; f < 4; ++f) dw[(t * 4 + o) * 4 + f] = fmaf(Tv[o], xw[f], dw[(t * 4 + o) * 4 + f]);
            }
        }
        if (gX) {
            *reinterpret_cast<float4*>(gX + (size_t)row * 4) =
                make_float4(gv[0] + old.x, gv[1] + old.y, gv[2] + old.z, gv[3] + old.w);
            if (accb) {
                const float xh[4] = {(xr.x - mu.x) * rs.x, (xr.y - mu.y) * rs.y, (xr.z - mu.z) * rs.z, (xr.w - mu.w) * rs.w};
#pragma unroll
                for (int f = 0; f < 4; ++f) { sg[f] = fmaf(w, gv[f], sg[f]); sgx[f] = fmaf(w * gv[f], xh[f], sgx[f]); }
            }
        }
    }
    flush_part<2>(dw, db, sg, sgx, red, sd, 12, false, accb);
}

__global__ void __launch_bounds__(THREADS, 1) mega_bwd_kernel(const __grid_constant__ Params P) {
    __shared__ __align__(16) float Wsm[MAX_SIDES * 80];   // per side [t][o][f] = W[o][t*4 + f], t < Cin / 4
    __shared__ __align__(16) float vec[48];               // gPre coefficients (12) | pad | self input (16) | cross input (16)
    __shared__ float red[(THREADS / 32) * 80];
    __shared__ int scratch[THREADS / 32];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < P.n_sides * 80; i += THREADS) {
        const int s = i / 80, j = i - s * 80;
        const Side& sd = P.s[s];
        const int c = (j >> 4) * 4 + (j & 3), o = (j >> 2) & 3;
        float v = 0.f;
        if (c < sd.Cin) v = o < sd.Ha ? sd.Wa[(size_t)o * sd.Cin + c] : sd.Wb[(size_t)(o - sd.Ha) * sd.Cin + c];
        Wsm[i] = v;
    }
    int *hdrN, *hdrE, *ent;
    stage_all(P, true, hdrN, hdrE, ent, scratch);
    unsigned target = 0;
    for (int s = P.n_sides - 1; s >= 0; --s) {
        const Side& sd = P.s[s];
        trace_mark(P, MAX_SIDES + s, 0);
        if (s < P.n_sides - 1) grid_sync(P.bar, target);
        trace_mark(P, MAX_SIDES + s, 1);
        const bool cross = sd.src_cross >= 0;
        const Tensor& to = P.t[sd.out];
        if (warp == 0) gpre_vectors(to, vec);
        else if (warp == 1) bn_vectors(P.t[sd.src_self], vec + 16);
        else if (warp == 2 && cross) bn_vectors(P.t[sd.src_cross], vec + 32);
        __syncthreads();
        trace_mark(P, MAX_SIDES + s, 2);
        Gpre gp;
        gp.c0 = lds4(vec); gp.c1 = lds4(vec + 4); gp.c2 = lds4(vec + 8);
        gp.relu_from = sd.relu_from;
        gp.bn = to.acc_b != nullptr;
        gp.need_z = gp.bn || sd.relu_from < 4;
        gp.G = to.grad;
        gp.Z = to.data;
        const float* Ws = Wsm + s * 80;
        if (sd.kind == 0) {
            bwd_self<0, 8>(P, sd, gp, Ws, vec + 16, hdrN, ent, red);
            if (cross) bwd_cross<0, 4>(P, sd, gp, Ws + 48, vec + 32, hdrE, ent, red);
        } else {
            bwd_self<1, 4>(P, sd, gp, Ws, vec + 16, hdrE, ent, red);
            if (cross) bwd_cross<1, 8>(P, sd, gp, Ws + 48, vec + 32, hdrN, ent, red);
        }
        trace_mark(P, MAX_SIDES + s, 3);
    }
    if (P.expand >= 0) {
        grid_sync(P.bar, target);
        expand_rows(P.g, P.t[P.expand].grad);
    }
    grid_exit(P.bar);
}

}  // namespace mk

// ---------------------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------------------
// profiling aid: device buffer of 2 * MAX_SIDES * grid * 4 time stamps (NULL = off); see profiles/mega_trace.py
static unsigned long long* g_trace = nullptr;
extern "C" int hgnn_mega_set_trace(void* dev_ptr) { g_trace = static_cast<unsigned long long*>(dev_ptr); return HGNN_OK; }
extern "C" int hgnn_mega_grid_for(int Rn, int n_act) {
    mk::Params p;
    p.g.Rn = Rn; p.g.n_act = n_act;
    extern int hgnn_mega_grid_of(const mk::Params&);
    return hgnn_mega_grid_of(p);
}
unsigned long long* hgnn_mega_trace_ptr(void) { return g_trace; }

static int mega_grid(const mk::Params& p) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = HGNN_SM_COUNT;
    }
    const long long most = p.g.Rn > p.g.n_act ? p.g.Rn : p.g.n_act;
    long long want = (most + 63) / 64;          // small batches: fewer CTAs, cheaper barriers
    if (want < 1) want = 1;
    return (int)(want < sms ? want : sms);
}

int hgnn_mega_grid_of(const mk::Params& p) { return mega_grid(p); }

// shared-memory budget of the structure cache: headers of the CTA's rows + as many entry words as fit
#define MEGA_DYN_SMEM_MAX (200 * 1024)
void hgnn_mega_plan(mk::Params* p) {
    p->grid = mega_grid(*p);
    p->max_n = (p->g.Rn + p->grid - 1) / p->grid + 1;
    p->max_e = p->g.n_act > 0 ? (p->g.n_act + p->grid - 1) / p->grid + 1 : 0;
    const long long hdr = (long long)p->max_n * mk::HDR_N + (long long)p->max_e * mk::HDR_E;
    long long cap = MEGA_DYN_SMEM_MAX / 4 - hdr;
    if (cap < 0) cap = 0;
    // no more than the structure can need: all entries of both row spaces spread over the grid, with slack
    const long long need = (2 * p->nnz1_n + 3 * p->nnz2 + 2 * p->nnz1_e + 3 * p->nnz2) / p->grid * 3 / 2 + 4096;
    if (p->nnz1_n >= 0 && need < cap) cap = need;
    p->cap_words = (int)cap;
}

static int mega_launch(const void* kernel, const mk::Params& p, cudaStream_t stream, const char* what) {
    if (!p.bar) {
        hgnn_set_error("%s: no barrier scratch (hgnn_batch_t.mega_scratch)", what);
        return HGNN_ERR_ARG;
    }
    static const void* attr_done[2] = {nullptr, nullptr};
    const size_t smem = ((size_t)p.max_n * mk::HDR_N + (size_t)p.max_e * mk::HDR_E + (size_t)p.cap_words) * sizeof(int);
    if (attr_done[0] != kernel && attr_done[1] != kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_DYN_SMEM_MAX);
        attr_done[attr_done[0] ? 1 : 0] = kernel;
    }
    void* args[] = {const_cast<mk::Params*>(&p)};
    cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(p.grid), dim3(mk::THREADS), args, smem, stream);
    if (e != cudaSuccess) {
        hgnn_set_error("%s: cudaLaunchCooperativeKernel: %s", what, cudaGetErrorString(e));
        return HGNN_ERR_CUDA;
    }
    return hgnn_check_launch(what);
}

int hgnn_mega_launch_fwd(const mk::Params& p, cudaStream_t stream) {
    return mega_launch(reinterpret_cast<const void*>(mk::mega_fwd_kernel), p, stream, "hgnn_mega_fwd");
}

int hgnn_mega_launch_bwd(const mk::Params& p, cudaStream_t stream) {
    return mega_launch(reinterpret_cast<const void*>(mk::mega_bwd_kernel), p, stream, "hgnn_mega_bwd");
}
