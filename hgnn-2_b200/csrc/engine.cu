// engine.cu -- the model-level LGNN / GNN training engine kernels ("v2" data flow).
//
// The layer-level kernels of side.cu mirror the reference's module boundaries: one launch for the
// fused side update, then separate launches for BN apply, BN backward reduce, ReLU/BN backward and
// the two transposed gathers - 12 launches per LGNN layer, each with its own ticket/finalise tail.
// At the script default h=2 a launch moves <= 25 MB, so the step is bound by launch count and by
// dependent memory hops, not by bandwidth (profiles/README.md).  The engine therefore keeps
// activations RAW (pre-batch-norm) in HBM and moves everything that used to be a launch into the
// prologue / gather of the neighbouring kernels:
//
//   forward  (1 launch per side):  consumers normalise their inputs on load (scale/shift derived in
//            the prologue from the producer's fp64 (sum z, sum z^2) accumulators), gather, concat,
//            conv, ReLU, write raw Z and add Z's own statistics to ITS accumulators.  No ticket.
//   backward (1 launch per side):  the BN + ReLU backward  gPre = (c0 g + c1 + c2 z) * mask  is
//            evaluated on the fly while gathering through the transposed operators; the launch
//            covers both the self rows and the cross rows, writes / accumulates the input
//            gradients, adds (sum g, sum g*xhat) of what it produced to the accumulators of the
//            tensors it differentiates, and adds dW / dbias partials to binned fp64 accumulators.
//   step end (1 launch):           hgnn_bins_reduce turns every accumulator into the flat fp32
//            gradient buffer; hgnn_bn_running_update applies the running-statistics rule.
//
// Reference semantics are unchanged (models/layers/layers_mnb.py:52-69,189-225,256-290,322-358,
// batch_normalization.py:34-43,65-93); parity is tested against the same golden vectors.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

#define ENG_THREADS 256
#define ENG_MAX_SMEM (200 * 1024)
#define ENG_LONG_ROW 32
#define ENG_CTA_ROW 1024
#define ENG_MAX_DEFER 192
#define ENG_BN_EPS 1e-5

namespace eng {

template <int VEC> struct V;
template <> struct V<1> {
    float v;
    __device__ __forceinline__ static V load(const float* p) { V r; r.v = __ldg(p); return r; }
    __device__ __forceinline__ static V loads(const float* p) { V r; r.v = *p; return r; }
    __device__ __forceinline__ static V zero() { V r; r.v = 0.f; return r; }
    __device__ __forceinline__ static V splat(float a) { V r; r.v = a; return r; }
    __device__ __forceinline__ void fma(float a, const V& x) { v = fmaf(a, x.v, v); }
    __device__ __forceinline__ void affine(const V& s, const V& t) { v = fmaf(v, s.v, t.v); }
    __device__ __forceinline__ void scale(float a) { v *= a; }
    __device__ __forceinline__ void add(const V& o) { v += o.v; }
    __device__ __forceinline__ void warp_reduce() { v = warp_sum(v); }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ void store_scalar(float* p) const { p[0] = v; }
    __device__ __forceinline__ float get(int) const { return v; }
    __device__ __forceinline__ void set(int, float a) { v = a; }
};
template <> struct V<4> {
    float4 v;
    __device__ __forceinline__ static V load(const float* p) { V r; r.v = __ldg(reinterpret_cast<const float4*>(p)); return r; }
    __device__ __forceinline__ static V loads(const float* p) { V r; r.v = *reinterpret_cast<const float4*>(p); return r; }
    __device__ __forceinline__ static V zero() { V r; r.v = make_float4(0.f, 0.f, 0.f, 0.f); return r; }
    __device__ __forceinline__ static V splat(float a) { V r; r.v = make_float4(a, a, a, a); return r; }
    __device__ __forceinline__ void fma(float a, const V& x) {
        v.x = fmaf(a, x.v.x, v.x); v.y = fmaf(a, x.v.y, v.y); v.z = fmaf(a, x.v.z, v.z); v.w = fmaf(a, x.v.w, v.w);
    }
    __device__ __forceinline__ void affine(const V& s, const V& t) {
        v.x = fmaf(v.x, s.v.x, t.v.x); v.y = fmaf(v.y, s.v.y, t.v.y);
        v.z = fmaf(v.z, s.v.z, t.v.z); v.w = fmaf(v.w, s.v.w, t.v.w);
    }
    __device__ __forceinline__ void scale(float a) { v.x *= a; v.y *= a; v.z *= a; v.w *= a; }
    __device__ __forceinline__ void add(const V& o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ __forceinline__ void warp_reduce() {
        v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
    }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void store_scalar(float* p) const { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
    __device__ __forceinline__ float get(int j) const { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }
    __device__ __forceinline__ void set(int j, float a) { if (j == 0) v.x = a; else if (j == 1) v.y = a; else if (j == 2) v.z = a; else v.w = a; }
};

// ---- row loaders -----------------------------------------------------------------------------
// forward: y = z * scale + shift  (the producer's batch-norm applied on load)
template <int VEC>
struct AffineLoader {
    const float* X;
    int ld;
    const float* sc;   // shared memory: scale[F]
    const float* sh;   // shared memory: shift[F]
    bool on;
    __device__ __forceinline__ V<VEC> operator()(int row, int xo) const {
        V<VEC> x = V<VEC>::load(X + (size_t)row * ld + xo);
        if (on) x.affine(V<VEC>::loads(sc + xo), V<VEC>::loads(sh + xo));
        return x;
    }
    // the loader is affine, so a weighted sum of loaded rows can be taken over the RAW rows and fixed up once:
    // sum_k v_k (x_k * sc + sh) = (sum_k v_k x_k) * sc + (sum_k v_k) * sh      (engine_wide.cuh)
    __device__ __forceinline__ V<VEC> raw(int row, int xo) const { return V<VEC>::load(X + (size_t)row * ld + xo); }
    __device__ __forceinline__ void finish(V<VEC>& acc, float vsum, int xo) const {
        if (!on) return;
        V<VEC> t = V<VEC>::loads(sh + xo);
        t.scale(vsum);
        acc.affine(V<VEC>::loads(sc + xo), t);
    }
};

// backward: gPre = (c0*g + c1 + c2*z) masked by the ReLU of the conv branch (z > 0 where f >= relu_from)
template <int VEC>
struct GpreLoader {
    const float* G;
    const float* Z;
    int ld;
    const float* c0;   // shared memory coefficient vectors [F]
    const float* c1;
    const float* c2;
    int relu_from;
    bool bn;
    __device__ __forceinline__ V<VEC> operator()(int row, int xo) const {
        V<VEC> g = V<VEC>::load(G + (size_t)row * ld + xo);
        if (!bn && relu_from >= ld) return g;
        V<VEC> z = V<VEC>::load(Z + (size_t)row * ld + xo);
        if (bn) {
            g.affine(V<VEC>::loads(c0 + xo), V<VEC>::loads(c1 + xo));
#pragma unroll
            for (int j = 0; j < VEC; ++j) g.set(j, fmaf(c2[xo + j], z.get(j), g.get(j)));
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            if (xo + j >= relu_from && !(z.get(j) > 0.f)) g.set(j, 0.f);
        return g;
    }
    // not linear in the stored rows (ReLU mask): the weighted sums are taken over the finished values
    __device__ __forceinline__ V<VEC> raw(int row, int xo) const { return (*this)(row, xo); }
    __device__ __forceinline__ void finish(V<VEC>&, float, int) const {}
};

// ---- gathers -----------------------------------------------------------------------------------
template <int VEC, typename L>
__device__ __forceinline__ V<VEC> gather_op(const OpList& ops, int t, int row, const L& ld_, int xo) {
    const int kind = ops.kind[t];
    if (kind == HGNN_OP_IDENT) return ld_(row, xo);
    if (kind == HGNN_OP_DIAG) {
        V<VEC> x = ld_(row, xo);
        x.scale(__ldg(ops.diag[t] + row));
        return x;
    }
    const int* __restrict__ col = ops.col[t];
    const float* __restrict__ val = ops.val[t];
    const int k0 = __ldg(ops.rowptr[t] + row), k1 = __ldg(ops.rowptr[t] + row + 1);
    V<VEC> acc = V<VEC>::zero();
    for (int k = k0; k < k1; k += 4) {           // batches of 4 independent gathers
        int c[4];
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool on = k + j < k1;
            c[j] = __ldg(col + (on ? k + j : k));
            v[j] = on ? __ldg(val + k + j) : 0.f;
        }
        V<VEC> x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ld_(c[j], xo);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.fma(v[j], x[j]);
    }
    return acc;
}

struct DeferList {
    int cnt;
    int rng_cnt;         // rows of the tile with a run-length part (at most one push per row)
    int rsum_id;         // range whose sum is cached in rsum (-1 = none)
    int items[ENG_MAX_DEFER];
    int rng_items[ENG_THREADS];
    float rsum[128];     // sum of the loader output over the cached row range, all features
};
__device__ __forceinline__ int defer_code(int t, int q, int r) { return (t << 24) | (q << 12) | r; }

template <int VEC, typename L>
__device__ __forceinline__ void gather_or_defer(const OpList& ops, int t, int row, const L& ld_, int xo,
                                                float* dst, DeferList* dl, int code) {
    if (ops.kind[t] == HGNN_OP_CSR) {
        // run-length part of the row (one push per row: the range sum covers every feature chunk)
        if (ops.rng_rowptr[t] && xo == 0 &&
            __ldg(ops.rng_rowptr[t] + row + 1) > __ldg(ops.rng_rowptr[t] + row)) {
            const int slot = atomicAdd(&dl->rng_cnt, 1);
            if (slot < ENG_THREADS) dl->rng_items[slot] = code;
            else __trap();       // more flagged (row, op) pairs than rows in a tile: impossible by construction
        }
        const int len = __ldg(ops.rowptr[t] + row + 1) - __ldg(ops.rowptr[t] + row);
        if (len > ENG_LONG_ROW) {
            const int slot = atomicAdd(&dl->cnt, 1);
            if (slot < ENG_MAX_DEFER) {
                dl->items[slot] = code;
                return;
            }
        }
    }
    gather_op<VEC>(ops, t, row, ld_, xo).store(dst);
}

template <int VEC, typename L>
__device__ __forceinline__ V<VEC> strided_gather(const OpList& ops, int t, int row, const L& ld_, int xo,
                                                 int first, int stride) {
    const int* __restrict__ col = ops.col[t];
    const float* __restrict__ val = ops.val[t];
    const int k0 = __ldg(ops.rowptr[t] + row), k1 = __ldg(ops.rowptr[t] + row + 1);
    V<VEC> a0 = V<VEC>::zero(), a1 = V<VEC>::zero();
    int k = k0 + first;
    for (; k + stride < k1; k += 2 * stride) {
        V<VEC> x0 = ld_(__ldg(col + k), xo);
        V<VEC> x1 = ld_(__ldg(col + k + stride), xo);
        a0.fma(__ldg(val + k), x0);
        a1.fma(__ldg(val + k + stride), x1);
    }
    if (k < k1) a0.fma(__ldg(val + k), ld_(__ldg(col + k), xo));
    a0.add(a1);
    return a0;
}

// all threads of the CTA; call between two __syncthreads()
template <int VEC, typename L>
__device__ __forceinline__ void gather_deferred(const OpList& ops, DeferList* dl, int row0, const L& ld_,
                                                int Fblk, float* tile, int Tp, float* wpart) {
    const int nd = min(dl->cnt, ENG_MAX_DEFER);
    const int nr = min(dl->rng_cnt, ENG_THREADS);
    if (nd == 0 && nr == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int it = warp; it < nd; it += nwarps) {
        const int code = dl->items[it];
        const int t = code >> 24, q = (code >> 12) & 0xfff, r = code & 0xfff;
        const int row = row0 + r, xo = q * VEC;
        const int len = __ldg(ops.rowptr[t] + row + 1) - __ldg(ops.rowptr[t] + row);
        if (len > ENG_CTA_ROW) continue;
        V<VEC> acc = strided_gather<VEC>(ops, t, row, ld_, xo, lane, 32);
        acc.warp_reduce();
        if (lane == 0) acc.store(tile + r * Tp + t * Fblk + xo);
    }
    for (int it = 0; it < nd; ++it) {
        const int code = dl->items[it];
        const int t = code >> 24, q = (code >> 12) & 0xfff, r = code & 0xfff;
        const int row = row0 + r, xo = q * VEC;
        const int len = __ldg(ops.rowptr[t] + row + 1) - __ldg(ops.rowptr[t] + row);
        if (len <= ENG_CTA_ROW) continue;
        V<VEC> acc = strided_gather<VEC>(ops, t, row, ld_, xo, threadIdx.x, blockDim.x);
        acc.warp_reduce();
        __syncthreads();
        if (lane == 0) acc.store_scalar(wpart + warp * 4);
        __syncthreads();
        if ((int)threadIdx.x < VEC) {
            float v = 0.f;
            for (int w = 0; w < nwarps; ++w) v += wpart[w * 4 + threadIdx.x];
            tile[r * Tp + t * Fblk + xo + threadIdx.x] = v;
        }
    }
    // ---- run-length parts: val * (sum of the loader output over a contiguous row range).  The
    //      range sum is computed once by the whole CTA and cached (the flagged rows of a graph are
    //      neighbours and share the phantom range of that graph).
    __syncthreads();
    for (int it = 0; it < nr; ++it) {
        const int code = dl->rng_items[it];
        const int t = code >> 24, r = code & 0xfff;
        const int row = row0 + r;
        const int e0 = __ldg(ops.rng_rowptr[t] + row), e1 = __ldg(ops.rng_rowptr[t] + row + 1);
        for (int e = e0; e < e1; ++e) {
            const int id = __ldg(ops.rng_id[t] + e);
            if (id != dl->rsum_id) {
                const int lo = __ldg(ops.rng_lo[t] + id), hi = __ldg(ops.rng_hi[t] + id);
                for (int q = 0; q < Fblk / VEC; ++q) {
                    V<VEC> a0 = V<VEC>::zero(), a1 = V<VEC>::zero();
                    int rr = lo + threadIdx.x;
                    for (; rr + (int)blockDim.x < hi; rr += 2 * blockDim.x) {
                        a0.add(ld_(rr, q * VEC));
                        a1.add(ld_(rr + blockDim.x, q * VEC));
                    }
                    if (rr < hi) a0.add(ld_(rr, q * VEC));
                    a0.add(a1);
                    a0.warp_reduce();
                    __syncthreads();
                    if (lane == 0) a0.store_scalar(wpart + warp * 4);
                    __syncthreads();
                    if ((int)threadIdx.x < VEC) {
                        float v = 0.f;
                        for (int w = 0; w < nwarps; ++w) v += wpart[w * 4 + threadIdx.x];
                        dl->rsum[q * VEC + threadIdx.x] = v;
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) dl->rsum_id = id;
                __syncthreads();
            }
            const float v = __ldg(ops.rng_val[t] + e);
            for (int f = threadIdx.x; f < Fblk; f += blockDim.x)
                tile[r * Tp + t * Fblk + f] = fmaf(v, dl->rsum[f], tile[r * Tp + t * Fblk + f]);
            __syncthreads();
        }
    }
}

// ---- accumulator bins ---------------------------------------------------------------------------
// out[c] = sum over bins of acc[b*width + c]; all threads; `scratch` = blockDim doubles.  Ends synced.
__device__ __forceinline__ void bins_total(const double* __restrict__ acc, int width, int nb,
                                           double* out, double* scratch) {
    if (nb * width <= 256) {                     // scratch holds 256 doubles: one parallel load round
        for (int e = threadIdx.x; e < nb * width; e += blockDim.x) scratch[e] = __ldcg(acc + e);
        __syncthreads();
        if ((int)threadIdx.x < width) {
            const int e = threadIdx.x;
            double t0 = 0.0, t1 = 0.0;
            for (int b = 0; b + 1 < nb; b += 2) {
                t0 += scratch[b * width + e];
                t1 += scratch[(b + 1) * width + e];
            }
            if (nb & 1) t0 += scratch[(nb - 1) * width + e];
            out[e] = t0 + t1;
        }
    } else {
        for (int c = threadIdx.x; c < width; c += blockDim.x) {
            double t = 0.0;
            for (int b = 0; b < nb; ++b) t += __ldcg(acc + (size_t)b * width + c);
            out[c] = t;
        }
    }
    __syncthreads();
}

// How to normalise a stored raw tensor on load.
struct BnRef {
    const double* acc;     // binned (sum z, sum z^2) of the producer, or NULL
    const float* affine;   // precomputed [scale(F), shift(F)] (eval mode), or NULL
    const float* w;        // scalar BN weight / bias (device)
    const float* b;
    int n;                 // rows behind the statistics
    double inv_n;          // 1 / n, divided on the host (an fp64 division is a ~40-instruction dependent sequence, and the
                           // coefficient round sits on the critical path of every launch)
};

// Fill scale/shift (and mean/rstd when asked) for a tensor of width F.  All threads; ends synced.
__device__ __forceinline__ bool bn_vectors(const BnRef& r, int F, float* sc, float* sh, float* mean,
                                           float* rstd, double* tot, double* scratch) {
    if (r.affine) {
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            sc[f] = r.affine[f];
            sh[f] = r.affine[F + f];
            if (mean) { mean[f] = 0.f; rstd[f] = 1.f; }
        }
        __syncthreads();
        return true;
    }
    if (!r.acc) {
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            sc[f] = 1.f;
            sh[f] = 0.f;
            if (mean) { mean[f] = 0.f; rstd[f] = 1.f; }
        }
        __syncthreads();
        return false;
    }
    bins_total(r.acc, 2 * F, hgnn_ws_bins(2 * F), tot, scratch);
    const double w = r.w[0], b = r.b[0];
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const double m = tot[f] / (double)r.n;
        double var = tot[F + f] / (double)r.n - m * m;
        if (var < 0.0) var = 0.0;
        const double sd = sqrt(var + ENG_BN_EPS);
        sc[f] = (float)(w / sd);
        sh[f] = (float)(b - w * m / sd);
        if (mean) { mean[f] = (float)m; rstd[f] = (float)(1.0 / sd); }
    }
    __syncthreads();
    return true;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct FwdArgs {
    int R;
    OpList ops;
    const float* Xs; int Fs; BnRef bn_s;
    const int* p_rowptr; const int* p_col; const float* p_pm; const float* p_pd;
    const float* Xc; int Fc; BnRef bn_c;
    const float* Wa; const float* ba; int Ha;
    const float* Wb; const float* bb; int Hb;
    int relu_from;
    float* Z;
    double* acc_out;     // binned (sum z, sum z^2) of Z, or NULL
    float* X1;           // row4 path only: save the concatenated x1 rows for the dW pass
    int TR, Cin, Cin_pad, Fout;
};

// WIDE (states of 32+ features, VEC = VOUT = 4): a warp's lanes walk the feature chunks of ONE row in the
// gather (coalesced 16 B x 32 lanes, CSR entries broadcast), and the linear is register-tiled 4 rows x 4
// outputs per thread (8 FMA per shared-memory load instead of 3.2).
template <int VEC, int VOUT, bool WIDE>
__global__ void __launch_bounds__(ENG_THREADS, WIDE ? 2 : 4)
fwd_kernel(const FwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double dscratch[ENG_THREADS];
    __shared__ double dtot[256];
    __shared__ DeferList dl;
    __shared__ float wpart[32];
    const int Cin = a.Cin, Cp = a.Cin_pad, Fout = a.Fout, TR = a.TR;
    const int K = a.ops.n, Fs = a.Fs, Fc = a.Fc;
    float* Wt = smem;                              // [Cin][Fout]
    float* bias = Wt + Cin * Fout;                 // [Fout]
    float* sc_s = bias + ((Fout + 3) & ~3);        // [Fs] scale / shift of the self input
    float* sh_s = sc_s + ((Fs + 3) & ~3);
    float* sc_c = sh_s + ((Fs + 3) & ~3);          // [Fc]
    float* sh_c = sc_c + ((Fc + 3) & ~3);
    float* tile = sh_c + ((Fc + 3) & ~3);          // [TR][Cp]
    const int tid = threadIdx.x;

    for (int i = tid; i < Cin * Fout; i += ENG_THREADS) {
        const int o = i / Cin, c = i - o * Cin;
        Wt[c * Fout + o] = (o < a.Ha) ? a.Wa[(size_t)o * Cin + c] : a.Wb[(size_t)(o - a.Ha) * Cin + c];
    }
    for (int o = tid; o < Fout; o += ENG_THREADS)
        bias[o] = (o < a.Ha) ? (a.ba ? a.ba[o] : 0.f) : (a.bb ? a.bb[o - a.Ha] : 0.f);
    if (tid == 0) dl.rsum_id = -1;
    const bool cross = a.p_rowptr != nullptr;
    const bool aff_s = bn_vectors(a.bn_s, Fs, sc_s, sh_s, nullptr, nullptr, dtot, dscratch);
    const bool aff_c = cross ? bn_vectors(a.bn_c, Fc, sc_c, sh_c, nullptr, nullptr, dtot, dscratch) : false;
    const AffineLoader<VEC> ls{a.Xs, Fs, sc_s, sh_s, aff_s};
    const AffineLoader<VEC> lc{a.Xc, Fc, sc_c, sh_c, aff_c};

    const int Qs = Fs / VEC, Qc = Fc / VEC;
    const int Q = Qs + (cross ? Qc : 0);
    const int xc0 = K * Fs;
    const int NQ = Fout / VOUT;
    const int rows_per_pass = ENG_THREADS / NQ;
    const bool owner = tid < rows_per_pass * NQ;
    const int oq = tid % NQ, rg = tid / NQ;
    float s1[VOUT], s2[VOUT];
#pragma unroll
    for (int j = 0; j < VOUT; ++j) s1[j] = s2[j] = 0.f;
    const int ntiles = (a.R + TR - 1) / TR;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, a.R - row0);
        if (tid == 0) { dl.cnt = 0; dl.rng_cnt = 0; }
        __syncthreads();
        for (int i = tid; i < Q * TR; i += ENG_THREADS) {
            const int q = WIDE ? i % Q : i / TR, r = WIDE ? i / Q : i - q * TR;
            if (r >= trc) continue;
            const int row = row0 + r;
            float* trow = tile + r * Cp;
            if (q < Qs) {
                const int xo = q * VEC;
                for (int t = 0; t < K; ++t)
                    gather_or_defer<VEC>(a.ops, t, row, ls, xo, trow + t * Fs + xo, &dl, defer_code(t, q, r));
            } else {
                const int xo = (q - Qs) * VEC;
                V<VEC> am = V<VEC>::zero(), ad = V<VEC>::zero();
                const int k0 = __ldg(a.p_rowptr + row), k1 = __ldg(a.p_rowptr + row + 1);
                for (int k = k0; k < k1; k += 4) {
                    int c[4];
                    float vm[4], vd[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool on = k + j < k1;
                        c[j] = __ldg(a.p_col + (on ? k + j : k));
                        vm[j] = on ? __ldg(a.p_pm + k + j) : 0.f;
                        vd[j] = on ? __ldg(a.p_pd + k + j) : 0.f;
                    }
                    V<VEC> x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = lc(c[j], xo);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        am.fma(vm[j], x[j]);
                        ad.fma(vd[j], x[j]);
                    }
                }
                am.store(trow + xc0 + xo);
                ad.store(trow + xc0 + Fc + xo);
            }
        }
        __syncthreads();
        gather_deferred<VEC>(a.ops, &dl, row0, ls, Fs, tile, Cp, wpart);
        __syncthreads();
        if (WIDE) {
            if (owner) {
                const float* w = Wt + oq * 4;
                const float4 b4 = *reinterpret_cast<const float4*>(bias + oq * 4);
                for (int rb = rg; rb < trc; rb += 4 * rows_per_pass) {
                    float acc[4][4];
                    const float* t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w;
                        const int r = rb + j * rows_per_pass;
                        t[j] = tile + (r < trc ? r : rb) * Cp;
                    }
                    for (int c = 0; c < Cin; c += 4) {
                        float x[4][4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 v = *reinterpret_cast<const float4*>(t[j] + c);
                            x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 wv = *reinterpret_cast<const float4*>(w + (c + u) * Fout);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[j][0] = fmaf(x[j][u], wv.x, acc[j][0]);
                                acc[j][1] = fmaf(x[j][u], wv.y, acc[j][1]);
                                acc[j][2] = fmaf(x[j][u], wv.z, acc[j][2]);
                                acc[j][3] = fmaf(x[j][u], wv.w, acc[j][3]);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = rb + j * rows_per_pass;
                        if (r >= trc) continue;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float v = acc[j][k];
                            if (oq * 4 + k >= a.relu_from) v = fmaxf(v, 0.f);
                            acc[j][k] = v;
                            s1[k % VOUT] += v;
                            s2[k % VOUT] = fmaf(v, v, s2[k % VOUT]);
                        }
                        *reinterpret_cast<float4*>(a.Z + (size_t)(row0 + r) * Fout + oq * 4) =
                            make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    }
                }
            }
        } else if (owner) {
            for (int r = rg; r < trc; r += rows_per_pass) {
                float acc[VOUT];
#pragma unroll
                for (int j = 0; j < VOUT; ++j) acc[j] = bias[oq * VOUT + j];
                const float* trow = tile + r * Cp;
                const float* w = Wt + oq * VOUT;
                if (VOUT == 4 && VEC == 4) {
                    for (int c = 0; c < Cin; c += 4) {
                        const float4 x = *reinterpret_cast<const float4*>(trow + c);
                        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 wv = *reinterpret_cast<const float4*>(w + (c + u) * Fout);
                            acc[0] = fmaf(xs[u], wv.x, acc[0]);
                            acc[1 % VOUT] = fmaf(xs[u], wv.y, acc[1 % VOUT]);
                            acc[2 % VOUT] = fmaf(xs[u], wv.z, acc[2 % VOUT]);
                            acc[3 % VOUT] = fmaf(xs[u], wv.w, acc[3 % VOUT]);
                        }
                    }
                } else {
#pragma unroll 4
                    for (int c = 0; c < Cin; ++c) {
                        const float x = trow[c];
#pragma unroll
                        for (int j = 0; j < VOUT; ++j) acc[j] = fmaf(x, w[c * Fout + j], acc[j]);
                    }
                }
                float* zrow = a.Z + (size_t)(row0 + r) * Fout + oq * VOUT;
#pragma unroll
                for (int j = 0; j < VOUT; ++j) {
                    float v = acc[j];
                    if (oq * VOUT + j >= a.relu_from) v = fmaxf(v, 0.f);
                    acc[j] = v;
                    s1[j] += v;
                    s2[j] = fmaf(v, v, s2[j]);
                }
                if (VOUT == 4) *reinterpret_cast<float4*>(zrow) = make_float4(acc[0], acc[1 % VOUT], acc[2 % VOUT], acc[3 % VOUT]);
                else zrow[0] = acc[0];
            }
        }
    }
    if (a.acc_out) {
        const int nb = hgnn_ws_bins(2 * Fout);
        const bool tree = (32 % NQ) == 0 && (ENG_THREADS % NQ) == 0;
#pragma unroll
        for (int j = 0; j < VOUT; ++j) {
            double x = 0.0, y = 0.0;
            if (tree) {
                x = cta_reduce_mod(owner ? (double)s1[j] : 0.0, NQ, dscratch);
                y = cta_reduce_mod(owner ? (double)s2[j] : 0.0, NQ, dscratch);
            } else {
                __syncthreads();
                dscratch[tid] = owner ? (double)s1[j] : 0.0;
                __syncthreads();
                if (tid < NQ) for (int k = 0; k < rows_per_pass; ++k) x += dscratch[k * NQ + tid];
                __syncthreads();
                dscratch[tid] = owner ? (double)s2[j] : 0.0;
                __syncthreads();
                if (tid < NQ) for (int k = 0; k < rows_per_pass; ++k) y += dscratch[k * NQ + tid];
            }
            if (tid < NQ) {
                accum_add(a.acc_out, 2 * Fout, nb, tid * VOUT + j, x);
                accum_add(a.acc_out, 2 * Fout, nb, Fout + tid * VOUT + j, y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward: one launch covers the self rows (tiles [0, tiles_self)) and the cross rows
// ---------------------------------------------------------------------------------------------
struct BwdPart {
    int R;                 // rows of the differentiated input (0 = part absent)
    OpList ops;            // transposed operators (cross: 2 CSR ops = Pm^T, Pd^T on one pattern)
    const float* X; int Fx; BnRef bn;      // the raw input and how it was normalised
    float* gX; int accumulate;             // gradient w.r.t. the NORMALISED input (NULL = not needed)
    double* acc_b;                         // binned (sum g, sum g*xhat) of that input's producer, or NULL
    int col0;                              // first weight column of this part
    int TR, nT, Tp, Xp, NG, P, tiles;
    size_t smem;
};

struct BwdArgs {
    // the side being differentiated
    const float* gY; const float* Z; int Fg; int relu_from; int Rg;
    const double* acc_f; const double* acc_b; const float* bn_w;   // acc_b == NULL: no batch-norm
    const float* Wa; int Ha; const float* Wb; int Hb; int Cin;
    double* dW_bins;       // [nb][Fg*Cin] (row o, column c) or NULL
    double* db_bins;       // [nb][Fg] or NULL
    BwdPart self, cross;
};

template <int VEC, int VOUT, bool WIDE>
__device__ __forceinline__ void bwd_part(const BwdArgs& a, const BwdPart& p, bool is_self, int first_tile,
                                         int tile_stride, float* smem, double* dscratch, double* dtot, DeferList* dl,
                                         float* wpart, const float* c0, const float* c1, const float* c2,
                                         bool has_bn) {
    const int nT = p.nT, Tp = p.Tp, Fx = p.Fx, Xp = p.Xp, Fg = a.Fg, TR = p.TR, P = p.P, NG = p.NG;
    // self tiles carry an extra own-row block (columns nT .. nT+Fg) whose column sums are dbias
    float* Wsm = smem;                                // [nT][Fx]
    float* sc = Wsm + ((nT * Fx + 3) & ~3);           // scale, shift, mean, rstd of the input [Fx] each
    float* sh = sc + ((Fx + 3) & ~3);
    float* mu = sh + ((Fx + 3) & ~3);
    float* rs = mu + ((Fx + 3) & ~3);
    float* tile = rs + ((Fx + 3) & ~3);               // [TR][Tp]
    float* xt = tile + TR * Tp;                       // [TR][Xp]   raw rows of the input
    float* dacc = xt + ((TR * Xp + 3) & ~3);          // [NG][P (+Fg)]
    const int PD = is_self ? P + Fg : P;
    const int tid = threadIdx.x;
    const int K = p.ops.n;
    const bool want_dw = a.dW_bins != nullptr;

    for (int i = tid; i < nT * Fx; i += ENG_THREADS) {
        const int c = i / Fx, f = i - c * Fx;
        const int t = c / Fg, o = c - t * Fg;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Wsm[i] = wrow[p.col0 + t * Fx + f];
    }
    for (int i = tid; i < NG * PD; i += ENG_THREADS) dacc[i] = 0.f;
    const bool x_aff = bn_vectors(p.bn, Fx, sc, sh, mu, rs, dtot, dscratch);
    const GpreLoader<VEC> lg{a.gY, a.Z, Fg, c0, c1, c2, a.relu_from, has_bn};

    const int Q = Fg / VEC;
    const int NQ = Fx / VOUT;
    const int rows_per_pass = ENG_THREADS / NQ;
    const bool owner = tid < rows_per_pass * NQ;
    const int fq = tid % NQ, rg = tid / NQ;
    const int PS = nT * NQ;
    float sg[VOUT], sgx[VOUT];
#pragma unroll
    for (int j = 0; j < VOUT; ++j) sg[j] = sgx[j] = 0.f;

    for (int tile_id = first_tile; tile_id < p.tiles; tile_id += tile_stride) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, p.R - row0);
        if (tid == 0) { dl->cnt = 0; dl->rng_cnt = 0; }
        __syncthreads();
        for (int i = tid; i < Q * TR; i += ENG_THREADS) {
            const int q = WIDE ? i % Q : i / TR, r = WIDE ? i / Q : i - q * TR;
            if (r >= trc) continue;
            const int xo = q * VEC;
            float* trow = tile + r * Tp;
            for (int t = 0; t < K; ++t)
                gather_or_defer<VEC>(p.ops, t, row0 + r, lg, xo, trow + t * Fg + xo, dl, defer_code(t, q, r));
            if (is_self) lg(row0 + r, xo).store(trow + nT + xo);     // own gPre row: dbias = column sums
        }
        if (VOUT == 4) {
            for (int i = tid; i < trc * NQ; i += ENG_THREADS) {
                const int r = i / NQ, g4 = i - r * NQ;
                *reinterpret_cast<float4*>(xt + r * Xp + g4 * 4) =
                    __ldg(reinterpret_cast<const float4*>(p.X + (size_t)row0 * Fx) + i);
            }
        } else {
            for (int i = tid; i < trc * Fx; i += ENG_THREADS) {
                const int r = i / Fx, f = i - r * Fx;
                xt[r * Xp + f] = p.X[(size_t)row0 * Fx + i];
            }
        }
        __syncthreads();
        gather_deferred<VEC>(p.ops, dl, row0, lg, Fg, tile, Tp, wpart);
        __syncthreads();
        // ---- gX = W^T T  (+ statistics of what was produced, for the input's own BN backward)
        if (WIDE) {
            if (p.gX && owner) {
                const float* w = Wsm + fq * 4;
                for (int rb = rg; rb < trc; rb += 4 * rows_per_pass) {
                    float acc[4][4];
                    const float* t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
                        const int r = rb + j * rows_per_pass;
                        t[j] = tile + (r < trc ? r : rb) * Tp;
                    }
                    for (int c = 0; c < nT; c += 4) {
                        float x[4][4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 v = *reinterpret_cast<const float4*>(t[j] + c);
                            x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 wv = *reinterpret_cast<const float4*>(w + (c + u) * Fx);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[j][0] = fmaf(x[j][u], wv.x, acc[j][0]);
                                acc[j][1] = fmaf(x[j][u], wv.y, acc[j][1]);
                                acc[j][2] = fmaf(x[j][u], wv.z, acc[j][2]);
                                acc[j][3] = fmaf(x[j][u], wv.w, acc[j][3]);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = rb + j * rows_per_pass;
                        if (r >= trc) continue;
                        if (p.acc_b) {
                            const float4 xr = *reinterpret_cast<const float4*>(xt + r * Xp + fq * 4);
                            const float xv[4] = {xr.x, xr.y, xr.z, xr.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int f = fq * 4 + k;
                                sg[k % VOUT] += acc[j][k];
                                sgx[k % VOUT] = fmaf(acc[j][k], (xv[k] - mu[f]) * rs[f], sgx[k % VOUT]);
                            }
                        }
                        float* dst = p.gX + (size_t)(row0 + r) * Fx + fq * 4;
                        float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                        if (p.accumulate) {
                            const float4 old = *reinterpret_cast<const float4*>(dst);
                            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                        }
                        *reinterpret_cast<float4*>(dst) = o;
                    }
                }
            }
        } else if (p.gX && owner) {
            for (int r = rg; r < trc; r += rows_per_pass) {
                float acc[VOUT];
#pragma unroll
                for (int j = 0; j < VOUT; ++j) acc[j] = 0.f;
                const float* trow = tile + r * Tp;
                const float* w = Wsm + fq * VOUT;
                if (VOUT == 4 && VEC == 4) {
                    for (int c = 0; c < nT; c += 4) {
                        const float4 x = *reinterpret_cast<const float4*>(trow + c);
                        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 wv = *reinterpret_cast<const float4*>(w + (c + u) * Fx);
                            acc[0] = fmaf(xs[u], wv.x, acc[0]);
                            acc[1 % VOUT] = fmaf(xs[u], wv.y, acc[1 % VOUT]);
                            acc[2 % VOUT] = fmaf(xs[u], wv.z, acc[2 % VOUT]);
                            acc[3 % VOUT] = fmaf(xs[u], wv.w, acc[3 % VOUT]);
                        }
                    }
                } else {
#pragma unroll 4
                    for (int c = 0; c < nT; ++c) {
                        const float x = trow[c];
#pragma unroll
                        for (int j = 0; j < VOUT; ++j) acc[j] = fmaf(x, w[c * Fx + j], acc[j]);
                    }
                }
                if (p.acc_b) {
#pragma unroll
                    for (int j = 0; j < VOUT; ++j) {
                        const int f = fq * VOUT + j;
                        const float xh = (xt[r * Xp + f] - mu[f]) * rs[f];
                        sg[j] += acc[j];
                        sgx[j] = fmaf(acc[j], xh, sgx[j]);
                    }
                }
                float* dst = p.gX + (size_t)(row0 + r) * Fx + fq * VOUT;
                if (VOUT == 4) {
                    float4 o = make_float4(acc[0], acc[1 % VOUT], acc[2 % VOUT], acc[3 % VOUT]);
                    if (p.accumulate) {
                        const float4 old = *reinterpret_cast<const float4*>(dst);
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    *reinterpret_cast<float4*>(dst) = o;
                } else {
                    dst[0] = p.accumulate ? dst[0] + acc[0] : acc[0];
                }
            }
        }
        // ---- dW[c][f] += sum_r T[r][c] * xnorm[r][f]  and (self) dbias[o] += sum_r gPre[r][o]
        if (want_dw && WIDE) {
            // 4x4 register blocks of dW (4 tile columns x 4 input features), rows strided over NG groups
            const int NB = (nT >> 2) * NQ;
            for (int sidx = tid; sidx < NG * NB; sidx += ENG_THREADS) {
                const int g = sidx / NB, b = sidx - g * NB;
                const int c4 = b / NQ, f4 = b - c4 * NQ;
                float4 scv = make_float4(1.f, 1.f, 1.f, 1.f), shv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (x_aff) {
                    scv = *reinterpret_cast<const float4*>(sc + f4 * 4);
                    shv = *reinterpret_cast<const float4*>(sh + f4 * 4);
                }
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                const float* tp = tile + c4 * 4;
                const float* xp = xt + f4 * 4;
#pragma unroll 2
                for (int r = g; r < trc; r += NG) {
                    const float4 tv = *reinterpret_cast<const float4*>(tp + r * Tp);
                    float4 xv = *reinterpret_cast<const float4*>(xp + r * Xp);
                    xv.x = fmaf(xv.x, scv.x, shv.x); xv.y = fmaf(xv.y, scv.y, shv.y);
                    xv.z = fmaf(xv.z, scv.z, shv.z); xv.w = fmaf(xv.w, scv.w, shv.w);
                    const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][0] = fmaf(tt[i], xv.x, acc[i][0]);
                        acc[i][1] = fmaf(tt[i], xv.y, acc[i][1]);
                        acc[i][2] = fmaf(tt[i], xv.z, acc[i][2]);
                        acc[i][3] = fmaf(tt[i], xv.w, acc[i][3]);
                    }
                }
                float* d = dacc + (size_t)g * PD + (c4 * 4) * Fx + f4 * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 o = *reinterpret_cast<float4*>(d + i * Fx);
                    o.x += acc[i][0]; o.y += acc[i][1]; o.z += acc[i][2]; o.w += acc[i][3];
                    *reinterpret_cast<float4*>(d + i * Fx) = o;
                }
            }
            if (is_self) {
                for (int sidx = tid; sidx < NG * Fg; sidx += ENG_THREADS) {
                    const int g = sidx / Fg, o = sidx - g * Fg;
                    float acc = 0.f;
                    for (int r = g; r < trc; r += NG) acc += tile[r * Tp + nT + o];
                    dacc[(size_t)g * PD + P + o] += acc;
                }
            }
        } else if (want_dw) {
            for (int sidx = tid; sidx < NG * PS; sidx += ENG_THREADS) {
                const int g = sidx / PS, q = sidx - g * PS;
                const int c = q / NQ, g4 = q - c * NQ;
                float acc[VOUT];
#pragma unroll
                for (int j = 0; j < VOUT; ++j) acc[j] = 0.f;
                float scv[VOUT], shv[VOUT];
#pragma unroll
                for (int j = 0; j < VOUT; ++j) {
                    scv[j] = x_aff ? sc[g4 * VOUT + j] : 1.f;
                    shv[j] = x_aff ? sh[g4 * VOUT + j] : 0.f;
                }
                for (int r = g; r < trc; r += NG) {
                    const float tv = tile[r * Tp + c];
#pragma unroll
                    for (int j = 0; j < VOUT; ++j)
                        acc[j] = fmaf(tv, fmaf(xt[r * Xp + g4 * VOUT + j], scv[j], shv[j]), acc[j]);
                }
                float* d = dacc + (size_t)g * PD + c * Fx + g4 * VOUT;
#pragma unroll
                for (int j = 0; j < VOUT; ++j) d[j] += acc[j];
            }
            if (is_self) {
                for (int sidx = tid; sidx < NG * Fg; sidx += ENG_THREADS) {
                    const int g = sidx / Fg, o = sidx - g * Fg;
                    float acc = 0.f;
                    for (int r = g; r < trc; r += NG) acc += tile[r * Tp + nT + o];
                    dacc[(size_t)g * PD + P + o] += acc;
                }
            }
        }
    }
    // ---- flush the CTA's partial sums to the binned fp64 accumulators
    __syncthreads();
    if (want_dw) {
        const int nbw = hgnn_ws_bins(Fg * a.Cin), nbb = hgnn_ws_bins(Fg);
        for (int q = tid; q < PD; q += ENG_THREADS) {
            float acc = 0.f;
            for (int g = 0; g < NG; ++g) acc += dacc[(size_t)g * PD + q];
            if (q < P) {
                const int c = q / Fx, f = q - c * Fx;
                const int t = c / Fg, o = c - t * Fg;
                accum_add(a.dW_bins, Fg * a.Cin, nbw, o * a.Cin + p.col0 + t * Fx + f, (double)acc);
            } else if (a.db_bins) {
                accum_add(a.db_bins, Fg, nbb, q - P, (double)acc);
            }
        }
    }
    if (p.acc_b && p.gX) {
        const int nb = hgnn_ws_bins(2 * Fx);
        const bool tree = (32 % NQ) == 0 && (ENG_THREADS % NQ) == 0;
#pragma unroll
        for (int j = 0; j < VOUT; ++j) {
            double x = 0.0, y = 0.0;
            if (tree) {
                x = cta_reduce_mod(owner ? (double)sg[j] : 0.0, NQ, dscratch);
                y = cta_reduce_mod(owner ? (double)sgx[j] : 0.0, NQ, dscratch);
            } else {
                __syncthreads();
                dscratch[tid] = owner ? (double)sg[j] : 0.0;
                __syncthreads();
                if (tid < NQ) for (int k = 0; k < rows_per_pass; ++k) x += dscratch[k * NQ + tid];
                __syncthreads();
                dscratch[tid] = owner ? (double)sgx[j] : 0.0;
                __syncthreads();
                if (tid < NQ) for (int k = 0; k < rows_per_pass; ++k) y += dscratch[k * NQ + tid];
            }
            if (tid < NQ) {
                accum_add(p.acc_b, 2 * Fx, nb, tid * VOUT + j, x);
                accum_add(p.acc_b, 2 * Fx, nb, Fx + tid * VOUT + j, y);
            }
        }
    }
}

template <int VEC, int VS, int VC, bool WIDE>
__global__ void __launch_bounds__(ENG_THREADS, WIDE ? 2 : 4)
bwd_kernel(const BwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double dscratch[ENG_THREADS];
    __shared__ double dtot[512];
    __shared__ DeferList dl;
    __shared__ float wpart[32];
    __shared__ __align__(16) float coef[3 * 128];
    const int Fg = a.Fg;
    if (threadIdx.x == 0) dl.rsum_id = -1;
    __syncthreads();
    float* c0 = coef;
    float* c1 = coef + 128;
    float* c2 = coef + 256;
    const bool has_bn = a.acc_b != nullptr;
    if (has_bn) {
        // coefficients of the BN backward of THIS side: gZ = c0 g + c1 + c2 z  (batch_normalization.py:65-77)
        double* tf = dtot;            // (sum z, sum z^2)
        double* tb = dtot + 2 * Fg;   // (sum g, sum g*xhat)
        bins_total(a.acc_f, 2 * Fg, hgnn_ws_bins(2 * Fg), tf, dscratch);
        bins_total(a.acc_b, 2 * Fg, hgnn_ws_bins(2 * Fg), tb, dscratch);
        const double w = a.bn_w[0], n = (double)a.Rg;
        for (int f = threadIdx.x; f < Fg; f += ENG_THREADS) {
            const double m = tf[f] / n;
            double var = tf[Fg + f] / n - m * m;
            if (var < 0.0) var = 0.0;
            const double sd = sqrt(var + ENG_BN_EPS);
            const double k0 = w / sd;
            const double k2 = -k0 * tb[Fg + f] / (n * sd);
            c0[f] = (float)k0;
            c2[f] = (float)k2;
            c1[f] = (float)(-k0 * tb[f] / n - k2 * m);
        }
        __syncthreads();
    }
    const int ts = a.self.R > 0 ? a.self.tiles : 0;
    // CTAs [0, ns) work on the self rows, the rest on the cross rows (both persistent over tiles)
    const int ns = a.self.R > 0 ? (a.cross.R > 0 ? max(1, (int)(((long long)gridDim.x * ts) / (ts + a.cross.tiles))) : gridDim.x) : 0;
    if ((int)blockIdx.x < ns) {
        bwd_part<VEC, VS, WIDE>(a, a.self, true, blockIdx.x, ns, smem, dscratch, dtot, &dl, wpart, c0, c1, c2, has_bn);
    } else {
        bwd_part<VEC, VC, WIDE>(a, a.cross, false, blockIdx.x - ns, gridDim.x - ns, smem, dscratch, dtot, &dl, wpart,
                          c0, c1, c2, has_bn);
    }
}

#include "engine_row4.cuh"
#include "engine_row4p.cuh"
#include "engine_rowg.cuh"
#include "engine_quad.cuh"
#include "engine_wide.cuh"
#include "engine_tc5.cuh"

}  // namespace eng

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline bool eng_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int eng_pad(int width, int vec) {
    if (vec == 4) {
        int p = (width + 3) & ~3;
        if (((p >> 2) & 1) == 0) p += 4;
        return p;
    }
    return width | 1;
}

// CTAs of `kernel` resident on the whole GPU with `smem` bytes of dynamic shared memory.  Cached per
// (kernel address, smem); also raises the kernel's opt-in shared-memory limit when needed.  (All
// instantiations of a kernel template share one function-pointer TYPE, so the cache must be keyed by
// the pointer VALUE.)
struct OccEntry { const void* fn; size_t smem; int occ; };
static int eng_resident_impl(const void* fn, size_t smem, int threads) {
    static OccEntry cache[64];
    static int n_cache = 0;
    static const void* attr_fn[32];
    static size_t attr_smem[32];
    static int n_attr = 0;
    if (smem > 24 * 1024) {      // static + dynamic above 48 KB needs the opt-in; static is < 10 KB here
        int i = 0;
        for (; i < n_attr; ++i) if (attr_fn[i] == fn) break;
        if (i == n_attr && n_attr < 32) { attr_fn[n_attr] = fn; attr_smem[n_attr] = 0; ++n_attr; }
        if (i < 32 && smem > attr_smem[i]) {
            cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_smem[i] = smem;
        }
    }
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].fn == fn && cache[i].smem == smem) return HGNN_SM_COUNT * cache[i].occ;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    if (n_cache < 64) { cache[n_cache].fn = fn; cache[n_cache].smem = smem; cache[n_cache].occ = occ; ++n_cache; }
    return HGNN_SM_COUNT * occ;
}
template <typename K>
static int eng_resident(K kernel, size_t smem) {
    return eng_resident_impl(reinterpret_cast<const void*>(kernel), smem, ENG_THREADS);
}

static eng::BnRef to_bnref(const hgnn_bn_ref_t* r) {
    eng::BnRef o;
    o.acc = r ? r->acc : nullptr;
    o.affine = r ? r->affine : nullptr;
    o.w = r ? r->weight : nullptr;
    o.b = r ? r->bias : nullptr;
    o.n = r ? r->n_rows : 0;
    o.inv_n = o.n > 0 ? 1.0 / (double)o.n : 0.0;
    return o;
}

// ---- programmatic dependent launch -----------------------------------------------------------------
static thread_local bool g_pdl = false;
void hgnn_eng_set_pdl(bool on) {
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("HGNN_B200_NO_PDL"); disabled = (e && e[0] == '1') ? 1 : 0; }
    g_pdl = on && !disabled;
}

// ---- launch recorder (csrc/program.cu: replay of a pass as ONE graph launch) --------------------------------
// While a recorder is installed on this thread, eng_launch does not launch: it appends (kernel, grid, block, shared
// memory, PDL flag, a copy of the by-value argument block) to the recorder.  The step executor then writes these into
// the kernel nodes of a graph it captured once (cudaGraphExecKernelNodeSetParams: 0.4 us per node against ~2 us for a
// launch, profiles/graph_update_probe.cu) and launches the graph.  Only kernels that go through eng_launch - the
// thread-per-row side kernels - can be recorded; hgnn_lg_side_fwd / _bwd refuse the other paths while recording.
static thread_local hgnn_eng_recorder_t* g_rec = nullptr;
void hgnn_eng_set_recorder(hgnn_eng_recorder_t* r) { g_rec = r; }
static inline bool eng_recording() { return g_rec != nullptr; }

template <typename Kernel, typename Args>
static void eng_launch(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t s, const Args& a) {
    if (g_rec) {
        hgnn_eng_slot_t slot;
        slot.func = reinterpret_cast<const void*>(kernel);
        slot.grid = grid; slot.block = threads; slot.smem = (unsigned)smem; slot.pdl = g_pdl ? 1 : 0;
        slot.args.assign(reinterpret_cast<const char*>(&a), reinterpret_cast<const char*>(&a) + sizeof(Args));
        g_rec->slots.push_back(std::move(slot));
        return;
    }
    if (!g_pdl) {
        kernel<<<grid, threads, smem, s>>>(a);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, a);
}

// ---- launch timeline (HGNN_B200_ABLATE bit 16): one g_ktrace slot per traced launch, in issue order -------------
static int g_ktrace_next = 0;
static int eng_trace_slot(int ablate) {
    if (!(ablate & 16) || g_ktrace_next >= KTRACE_SLOTS) return -1;
    return g_ktrace_next++;
}

// ---- thread-per-row fast path for width-4 states (h = 2) ---------------------------------------
static bool eng_row4_ops(const hgnn_op_t* ops, int n_ops) {
    if (n_ops < 3 || n_ops > 4) return false;
    if (ops[0].kind != HGNN_OP_IDENT || ops[1].kind != HGNN_OP_DIAG) return false;
    for (int i = 2; i < n_ops; ++i)
        if (ops[i].kind != HGNN_OP_CSR) return false;
    return true;
}

static bool eng_row4_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_NO_ROW4"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

extern "C" int hgnn_lg_row4_eligible(const hgnn_op_t* ops, int n_ops, int Fs, int Fc, int Fout) {
    if (eng_row4_disabled()) return 0;
    if (Fs != 4 || Fout != 4 || (Fc != 0 && Fc != 4) || !ops || !eng_row4_ops(ops, n_ops)) return 0;
    for (int i = 2; i < n_ops; ++i) if (ops[i].rng_rowptr) return 0;
    return 1;
}

static bool eng_try_fwd_row4(const eng::FwdArgs& g, const hgnn_side_t* side, hgnn_stream_t stream) {
    const bool cross = g.p_rowptr != nullptr;
    if (!hgnn_lg_row4_eligible(side->ops, side->n_ops, g.Fs, cross ? g.Fc : 0, g.Fout)) return false;
    if (!eng_aligned16(g.Xs) || !eng_aligned16(g.Z) || (cross && !eng_aligned16(g.Xc)) ||
        (g.X1 && !eng_aligned16(g.X1))) return false;
    eng::Fwd4Args a;
    a.R = g.R; a.n_csr = side->n_ops - 2; a.diag = side->ops[1].diag;
    for (int i = 0; i < 2; ++i) {
        const bool on = i < a.n_csr;
        a.rowptr[i] = on ? side->ops[2 + i].rowptr : nullptr;
        a.col[i] = on ? side->ops[2 + i].col : nullptr;
        a.val[i] = on ? side->ops[2 + i].val : nullptr;
    }
    a.Xs = g.Xs; a.bn_s = g.bn_s;
    a.p_rowptr = g.p_rowptr; a.p_col = g.p_col; a.p_pm = g.p_pm; a.p_pd = g.p_pd; a.Xc = g.Xc; a.bn_c = g.bn_c;
    a.Wa = g.Wa; a.ba = g.ba; a.Ha = g.Ha; a.Wb = g.Wb; a.bb = g.bb; a.Hb = g.Hb;
    a.relu_from = g.relu_from; a.Cin = g.Cin; a.Z = g.Z; a.acc_out = g.acc_out; a.X1 = g.X1;
    a.roww = side->roww; a.rowmap = side->rowmap;
    static int ablate = -1;
    if (ablate < 0) { const char* e = getenv("HGNN_B200_ABLATE"); ablate = e ? atoi(e) : 0; }
    a.ablate = ablate & ~16;
    a.trace_slot = eng_trace_slot(ablate);
    cudaStream_t s = to_stream(stream);
    // several lanes per row (engine_quad.cuh): one CSR operator, no saved x1 rows.  Opt-in (HGNN_B200_QUAD=1;
    // HGNN_B200_QUAD_LPR=<heavy>,<light> picks the lanes per row): measured SLOWER on C2 - node side 11.4 vs 7.6 us,
    // edge side 16.0 vs 7.2 us with 4 lanes, 10.6 / 10.6 us with 2 (profiles/logs/bench_r2x_quad*.log).  ncu
    // (profiles/prof_quad_r2y_raw.csv): 2.7x the warp instructions of the thread-per-row kernel (4.2 M vs 1.6 M: every
    // lane repeats the structure loads, address arithmetic and the batch-norm prologue) at 1.8x its issue rate.
    static int quad = -1, lpr_heavy = 4, lpr_light = 2;
    if (quad < 0) {
        const char* e = getenv("HGNN_B200_QUAD");
        quad = (e && e[0] == '1') ? 1 : 0;
        const char* l = getenv("HGNN_B200_QUAD_LPR");
        if (l) { lpr_heavy = atoi(l); const char* c = strchr(l, ','); lpr_light = c ? atoi(c + 1) : lpr_heavy; }
    }
    if (quad && a.n_csr == 1 && !a.X1 && !a.ablate) {
        const double per_row = a.R > 0 ? ((double)side->ops[2].nnz + (cross ? (double)side->p_nnz : 0.0)) / a.R : 0.0;
        const int lpr = per_row > 8.0 ? lpr_heavy : lpr_light;
#define QD_FWD(CROSS, LPR, JA, JP)                                                                                   \
        {                                                                                                            \
            const int rpc = QD_THREADS / LPR;                                                                        \
            const int cap = eng_resident_impl((const void*)eng::fwd_quad_kernel<CROSS, LPR, JA, JP>, 0, QD_THREADS); \
            const int grid = min(ceil_div(a.R, rpc), cap);                                                           \
            eng_launch(eng::fwd_quad_kernel<CROSS, LPR, JA, JP>, grid, QD_THREADS, 0, s, a);                         \
        }
        if (lpr == 4) { if (cross) QD_FWD(true, 4, 2, 4) else QD_FWD(false, 4, 2, 1) return true; }
        if (lpr == 2) { if (cross) QD_FWD(true, 2, 2, 2) else QD_FWD(false, 2, 4, 1) return true; }
#undef QD_FWD
    }
    const int want = ceil_div(a.R, R4_THREADS);
    // entries per gather batch from the average row length (nnz hints; unknown -> 4): a typical row
    // should fit ONE batch so that its loads form three dependent rounds in total
    const double avg_a = a.R > 0 ? (double)side->ops[2].nnz / a.R : 0.0;
    const double avg_p = (cross && a.R > 0) ? (double)side->p_nnz / a.R : 0.0;
    bool big_a = avg_a > 3.5, big_p = avg_p > 5.0;
    static int full_grid = -1;
    if (full_grid < 0) { const char* e = getenv("HGNN_B200_FWD_FULLGRID"); full_grid = (e && e[0] == '1') ? 1 : 0; }
    static int force = -2;
    if (force == -2) { const char* e = getenv("HGNN_B200_FWD_BATCH"); force = e ? atoi(e) : -1; }   // bit 0: big_a, bit 1: big_p
    if (force >= 0) { big_a = force & 1; big_p = (force & 2) != 0; }
#define R4_FWD(NCSR, CROSS, BA, BP)                                                                      \
    {                                                                                                    \
        const int cap_ = eng_resident_impl((const void*)eng::fwd_row4_kernel<NCSR, CROSS, BA, BP>, 0, R4_THREADS);   \
        int grid = full_grid ? want : min(want, cap_);                                                              \
        eng_launch(eng::fwd_row4_kernel<NCSR, CROSS, BA, BP>, grid, R4_THREADS, 0, s, a);                \
    }
#define R4_FWD_B(NCSR)                                                                                   \
    if (!cross) { if (big_a) R4_FWD(NCSR, false, 8, 4) else R4_FWD(NCSR, false, 4, 4) }                  \
    else if (big_a) { if (big_p) R4_FWD(NCSR, true, 8, 16) else R4_FWD(NCSR, true, 8, 4) }               \
    else { if (big_p) R4_FWD(NCSR, true, 4, 16) else R4_FWD(NCSR, true, 4, 4) }
    if (a.n_csr == 1) { R4_FWD_B(1) } else { R4_FWD_B(2) }
#undef R4_FWD_B
#undef R4_FWD
    return true;
}

// thread-per-row kernels for the odd small widths of layer 0 and the readout (engine_rowg.cuh)
static bool eng_try_fwd_rowg(const eng::FwdArgs& g, const hgnn_side_t* side, hgnn_stream_t stream) {
    if (eng_row4_disabled() || g.X1) return false;
    const bool cross = g.p_rowptr != nullptr;
    const int Fs = g.Fs, Fc = cross ? g.Fc : 0, Fo = g.Fout;
    if (!side->ops || !eng_row4_ops(side->ops, side->n_ops)) return false;
    for (int i = 2; i < side->n_ops; ++i) if (side->ops[i].rng_rowptr) return false;
    if (Fo < 1 || Fo > 4 || (g.acc_out && Fo != 4)) return false;
    if ((g.bn_s.acc || g.bn_s.affine) && Fs != 4) return false;
    if (cross && (g.bn_c.acc || g.bn_c.affine) && Fc != 4) return false;
    if ((Fs == 4 && !eng_aligned16(g.Xs)) || (Fc == 4 && !eng_aligned16(g.Xc)) || (Fo == 4 && !eng_aligned16(g.Z))) return false;
    eng::Fwd4Args a;
    a.R = g.R; a.n_csr = side->n_ops - 2; a.diag = side->ops[1].diag;
    for (int i = 0; i < 2; ++i) {
        const bool on = i < a.n_csr;
        a.rowptr[i] = on ? side->ops[2 + i].rowptr : nullptr;
        a.col[i] = on ? side->ops[2 + i].col : nullptr;
        a.val[i] = on ? side->ops[2 + i].val : nullptr;
    }
    a.Xs = g.Xs; a.bn_s = g.bn_s;
    a.p_rowptr = g.p_rowptr; a.p_col = g.p_col; a.p_pm = g.p_pm; a.p_pd = g.p_pd; a.Xc = g.Xc; a.bn_c = g.bn_c;
    a.Wa = g.Wa; a.ba = g.ba; a.Ha = g.Ha; a.Wb = g.Wb; a.bb = g.bb; a.Hb = g.Hb;
    a.relu_from = g.relu_from; a.Cin = g.Cin; a.Z = g.Z; a.acc_out = g.acc_out; a.X1 = nullptr; a.ablate = 0; a.trace_slot = -1;
    a.roww = side->roww; a.rowmap = side->rowmap;
    cudaStream_t s = to_stream(stream);
    const int want = ceil_div(a.R, R4_THREADS);
    const double avg_a = a.R > 0 ? (double)side->ops[2].nnz / a.R : 0.0;
    const double avg_p = (cross && a.R > 0) ? (double)side->p_nnz / a.R : 0.0;
    const bool big = avg_a > 3.5 || avg_p > 5.0;       // node-like rows: batches of 8 / 16, else 4 / 4
    bool done = false;
#define RG_FWD(NCSR, CROSS, FS, FC, FO)                                                                             \
    if (!done && a.n_csr == NCSR && cross == CROSS && Fs == FS && Fc == (CROSS ? FC : 0) && Fo == FO) {             \
        if (big) {                                                                                                  \
            int grid = min(want, eng_resident_impl((const void*)eng::fwd_rowg_kernel<NCSR, CROSS, 8, 16, FS, FC, FO>, 0, R4_THREADS)); \
            eng_launch(eng::fwd_rowg_kernel<NCSR, CROSS, 8, 16, FS, FC, FO>, grid, R4_THREADS, 0, s, a);            \
        } else {                                                                                                    \
            int grid = min(want, eng_resident_impl((const void*)eng::fwd_rowg_kernel<NCSR, CROSS, 4, 4, FS, FC, FO>, 0, R4_THREADS)); \
            eng_launch(eng::fwd_rowg_kernel<NCSR, CROSS, 4, 4, FS, FC, FO>, grid, R4_THREADS, 0, s, a);             \
        }                                                                                                           \
        done = true;                                                                                                \
    }
    // LGNN (models/gnns/model_mnb.py:98-100): layer 0 node side [5 | 1 -> 4], edge side [1 | 4 -> 4], readout [4 | 4 -> 1, 2]
    RG_FWD(1, true, 5, 1, 4) RG_FWD(1, true, 1, 4, 4) RG_FWD(1, true, 4, 4, 2) RG_FWD(1, true, 4, 4, 1)
    RG_FWD(2, true, 5, 1, 4) RG_FWD(2, true, 1, 4, 4) RG_FWD(2, true, 4, 4, 2) RG_FWD(2, true, 4, 4, 1)
    // power GNN (:48-50): layer 0 [5 -> 4], readout [4 -> 1, 2]
    RG_FWD(1, false, 5, 0, 4) RG_FWD(1, false, 4, 0, 2) RG_FWD(1, false, 4, 0, 1)
    RG_FWD(2, false, 5, 0, 4) RG_FWD(2, false, 4, 0, 2) RG_FWD(2, false, 4, 0, 1)
#undef RG_FWD
    return done;
}

static bool eng_part_vout4(const float* X, const float* gX, int Fx) {
    return (Fx % 4 == 0) && eng_aligned16(X) && (!gX || eng_aligned16(gX));
}

// ---- tensor-core tile kernels for wide states (engine_wide.cuh) -----------------------------------
#define WD_MAX_SMEM (212 * 1024)     // dynamic; the kernels hold <= 13 KB of static shared memory on top
static bool eng_wide_disabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_NO_WIDE_MMA"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
// smallest output width that takes the tensor-core kernels (HGNN_B200_WIDE_MIN overrides; a multiple of 16)
static int eng_wide_min() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_WIDE_MIN"); v = e ? atoi(e) : 32; if (v < 16) v = 16; }
    return v;
}

// width-only part of the forward dispatch: tile rows and dynamic shared memory, or false
static bool eng_wide_fwd_fits(int n_ops, int Fs, int Fc, int Fout, int* TR_out, size_t* smem_out) {
    if (eng_wide_disabled()) return false;
    if (Fout % 16 || Fout < eng_wide_min() || Fout > 128) return false;
    if (Fs < 8 || Fs % 8 || Fc % 8 || Fs > 128 || Fc > 128) return false;   // two 16-byte chunks per gather item
    const int Cin = n_ops * Fs + 2 * Fc;
    const int Cp = eng_pad(Cin, 4);
    for (int TR = WD_TR; TR >= 16; TR >>= 1) {
        const size_t smem = (size_t)eng::wide_fwd_layout(Cin, Cp, Fout, Fs, Fc, TR).total * sizeof(float);
        if (smem <= WD_MAX_SMEM) {
            *TR_out = TR;
            *smem_out = smem;
            return true;
        }
    }
    return false;
}

static bool eng_try_fwd_wide(eng::FwdArgs& a, hgnn_stream_t stream) {
    int TR = 0;
    size_t smem = 0;
    if (!eng_wide_fwd_fits(a.ops.n, a.Fs, a.Fc, a.Fout, &TR, &smem)) return false;
    if (!eng_aligned16(a.Xs) || (a.Fc && !eng_aligned16(a.Xc)) || !eng_aligned16(a.Z)) return false;
    a.Cin_pad = eng_pad(a.Cin, 4);
    a.TR = TR;
    const int ntiles = ceil_div(a.R, TR);
    const int grid = balanced_grid(ntiles, eng_resident_impl(reinterpret_cast<const void*>(eng::fwd_wide_kernel), smem, WD_THREADS));
    eng::fwd_wide_kernel<<<grid, WD_THREADS, smem, to_stream(stream)>>>(a);
    return true;
}

// EXPERIMENTAL tcgen05 forward (engine_tc5.cuh): only with HGNN_B200_WIDE_TC5=1
static bool eng_tc5_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_WIDE_TC5"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
static bool eng_tc5_bwd_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_WIDE_TC5_BWD"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
static bool eng_try_fwd_tc5(eng::FwdArgs& a, hgnn_stream_t stream) {
    if (!eng_tc5_enabled() || eng_wide_disabled()) return false;
    const bool cross = a.p_rowptr != nullptr;
    if (a.Fs != 32 && a.Fs != 64) return false;
    if (cross && a.Fc != 32 && a.Fc != 64) return false;
    if (a.Fout != 32 && a.Fout != 64) return false;
    if ((a.ops.n + (cross ? 2 : 0) + 2) * a.Fout > 512) return false;        // TMEM columns
    for (int t = 0; t < a.ops.n; ++t) if (a.ops.rng_rowptr[t]) return false;
    if (!eng_aligned16(a.Xs) || (cross && !eng_aligned16(a.Xc)) || !eng_aligned16(a.Z)) return false;
    const size_t smem = (size_t)eng::tc5_layout(a.Cin, a.Fout, a.Fs, a.Fc).total * sizeof(float);
    if (smem > WD_MAX_SMEM) return false;
    a.TR = 64;
    const int ntiles = ceil_div(a.R, 64);
    const int grid = balanced_grid(ntiles, eng_resident_impl(reinterpret_cast<const void*>(eng::fwd_tc5_kernel), smem, WD_THREADS));
    eng::fwd_tc5_kernel<<<grid, WD_THREADS, smem, to_stream(stream)>>>(a);
    return true;
}

// width-only part of the backward dispatch for one part (n_ops operators gathered, gPre width Fg, input width Fx)
static bool eng_wide_part_fits(int n_ops, int Fg, int Fx, bool is_self, int* TR_out, size_t* smem_out) {
    if (Fx % 16 || Fx > 128) return false;
    const int nT = n_ops * Fg;
    const int njx = Fx / 16;                 // dW: a warp owns one 16-column block and <= WD_MAXDW row blocks
    if (WD_WARPS % njx || nT / 16 > WD_MAXDW * (WD_WARPS / njx)) return false;
    const int Tp = eng_pad(is_self ? nT + Fg : nT, 4);
    for (int TR = WD_TR; TR >= 16; TR >>= 1) {
        const size_t smem = (size_t)eng::wide_bwd_layout(nT, Fx, Tp, Fx + 8, TR).total * sizeof(float);
        if (smem <= WD_MAX_SMEM) {
            *TR_out = TR;
            *smem_out = smem;
            return true;
        }
    }
    return false;
}
static bool eng_wide_bwd_width(int Fg) {
    return !eng_wide_disabled() && Fg % 16 == 0 && Fg >= eng_wide_min() && Fg <= 128;
}

// derived fields of one backward part for bwd_wide_kernel; false if it does not qualify / fit
static bool eng_plan_part_wide(eng::BwdPart& p, int Fg, bool is_self) {
    int TR = 0;
    size_t smem = 0;
    if (!eng_part_vout4(p.X, p.gX, p.Fx) || !eng_wide_part_fits(p.ops.n, Fg, p.Fx, is_self, &TR, &smem)) return false;
    p.nT = p.ops.n * Fg;
    p.P = p.nT * p.Fx;
    p.NG = 1;
    p.Tp = eng_pad(is_self ? p.nT + Fg : p.nT, 4);
    p.Xp = p.Fx + 8;
    p.TR = TR;
    p.smem = smem;
    p.tiles = ceil_div(p.R, TR);
    return true;
}

extern "C" int hgnn_lg_wide_eligible(int n_ops, int Fs, int Fc, int Fout, int backward) {
    int TR = 0;
    size_t smem = 0;
    if (n_ops < 1 || n_ops > HGNN_MAX_OPS || Fs < 1 || Fc < 0 || Fout < 1) return 0;
    if (!backward) return eng_wide_fwd_fits(n_ops, Fs, Fc, Fout, &TR, &smem) ? 1 : 0;
    if (!eng_wide_bwd_width(Fout)) return 0;
    if (!eng_wide_part_fits(n_ops, Fout, Fs, true, &TR, &smem)) return 0;
    if (Fc > 0 && !eng_wide_part_fits(2, Fout, Fc, false, &TR, &smem)) return 0;
    return 1;
}

extern "C" int hgnn_lg_side_fwd(const hgnn_side_t* side, const hgnn_bn_ref_t* bn_self,
                                const hgnn_bn_ref_t* bn_cross, const float* Wa, const float* ba, int Ha,
                                const float* Wb, const float* bb, int Hb, int relu_from, float* Z,
                                double* acc_out, float* X1, hgnn_stream_t stream) {
    HGNN_REQUIRE(side && Z, "null argument");
    eng::FwdArgs a;
    a.X1 = X1;
    HGNN_REQUIRE(make_oplist(side->ops, side->n_ops, &a.ops) == 0 && side->n_ops >= 1, "bad operator list");
    a.R = side->R;
    a.Xs = side->Xs; a.Fs = side->Fs; a.bn_s = to_bnref(bn_self);
    a.p_rowptr = side->p_rowptr; a.p_col = side->p_col; a.p_pm = side->p_pm; a.p_pd = side->p_pd;
    a.Xc = side->Xc; a.Fc = side->p_rowptr ? side->Fc : 0; a.bn_c = to_bnref(bn_cross);
    a.Wa = Wa; a.ba = ba; a.Ha = Ha; a.Wb = Wb; a.bb = bb; a.Hb = Hb;
    a.relu_from = relu_from; a.Z = Z; a.acc_out = acc_out;
    HGNN_REQUIRE(a.R >= 0 && a.Fs >= 1 && a.Xs, "bad self features");
    HGNN_REQUIRE(!side->p_rowptr || (side->Fc >= 1 && side->Xc && side->p_col && side->p_pm && side->p_pd), "bad cross part");
    HGNN_REQUIRE(Ha >= 0 && Hb >= 0 && Ha + Hb >= 1 && Ha + Hb <= 128, "output width must be in [1, 128]");
    HGNN_REQUIRE(a.Fs <= 128 && a.Fc <= 128, "input widths must be <= 128");
    HGNN_REQUIRE((Ha == 0 || Wa) && (Hb == 0 || Wb), "null weights");
    HGNN_REQUIRE(!(a.bn_s.acc) || (a.bn_s.w && a.bn_s.b && a.bn_s.n > 0), "bad self bn reference");
    HGNN_REQUIRE(!(a.bn_c.acc) || (a.bn_c.w && a.bn_c.b && a.bn_c.n > 0), "bad cross bn reference");
    if (a.R == 0) return HGNN_OK;
    a.Fout = Ha + Hb;
    a.Cin = side->n_ops * a.Fs + 2 * a.Fc;
    if (eng_try_fwd_row4(a, side, stream)) return hgnn_check_launch("hgnn_lg_side_fwd(row4)");
    if (eng_try_fwd_rowg(a, side, stream)) return hgnn_check_launch("hgnn_lg_side_fwd(rowg)");
    HGNN_REQUIRE(!eng_recording(), "recording: this side does not run on the thread-per-row kernels");
    HGNN_REQUIRE(!X1, "x1 rows can only be saved by the width-4 fast path (check hgnn_lg_row4_eligible)");
    HGNN_REQUIRE(!side->roww && !side->rowmap, "row weights need the thread-per-row kernels (widths of the h = 2 feature maps)");
    if (eng_try_fwd_tc5(a, stream)) return hgnn_check_launch("hgnn_lg_side_fwd(tc5)");
    if (eng_try_fwd_wide(a, stream)) return hgnn_check_launch("hgnn_lg_side_fwd(wide)");
    const bool vec4 = (a.Fs % 4 == 0) && (a.Fc % 4 == 0) && eng_aligned16(a.Xs) && (a.Fc == 0 || eng_aligned16(a.Xc));
    const bool vout4 = vec4 && (a.Fout % 4 == 0) && eng_aligned16(Z);
    a.Cin_pad = eng_pad(a.Cin, vec4 ? 4 : 1);
    const int rows_per_pass = ENG_THREADS / (vout4 ? a.Fout / 4 : a.Fout);
    const int items_per_row = (a.Fs + a.Fc) / (vec4 ? 4 : 1);
    // register-tiled variant from 32 outputs up: measured at h = 8 (16 outputs) it is slower (10.3 vs 7.6 ms per step),
    // at h = 32 faster (48.9 vs 84.6 ms)
    const bool wide = vec4 && vout4 && a.Fout >= 32 && items_per_row >= 8;
    int TR = wide ? 4 * rows_per_pass : min(rows_per_pass, max(32, ENG_THREADS / max(1, items_per_row)));
    size_t fixed = ((size_t)a.Cin * a.Fout + ((a.Fout + 3) & ~3) + 2 * ((a.Fs + 3) & ~3) + 2 * ((a.Fc + 3) & ~3)) * sizeof(float);
    while (TR > 1 && fixed + (size_t)TR * a.Cin_pad * sizeof(float) > ENG_MAX_SMEM) TR >>= 1;
    size_t smem = fixed + (size_t)TR * a.Cin_pad * sizeof(float);
    if (smem > ENG_MAX_SMEM) {
        hgnn_set_error("hgnn_lg_side_fwd: Cin=%d x Fout=%d does not fit shared memory", a.Cin, a.Fout);
        return HGNN_ERR_ARG;
    }
    a.TR = TR;
    const int ntiles = ceil_div(a.R, TR);
    cudaStream_t s = to_stream(stream);
    if (wide) {
        int grid = balanced_grid(ntiles, eng_resident(eng::fwd_kernel<4, 4, true>, smem));
        eng::fwd_kernel<4, 4, true><<<grid, ENG_THREADS, smem, s>>>(a);
    } else if (vec4 && vout4) {
        int grid = balanced_grid(ntiles, eng_resident(eng::fwd_kernel<4, 4, false>, smem));
        eng::fwd_kernel<4, 4, false><<<grid, ENG_THREADS, smem, s>>>(a);
    } else if (vec4) {
        int grid = balanced_grid(ntiles, eng_resident(eng::fwd_kernel<4, 1, false>, smem));
        eng::fwd_kernel<4, 1, false><<<grid, ENG_THREADS, smem, s>>>(a);
    } else {
        int grid = balanced_grid(ntiles, eng_resident(eng::fwd_kernel<1, 1, false>, smem));
        eng::fwd_kernel<1, 1, false><<<grid, ENG_THREADS, smem, s>>>(a);
    }
    return hgnn_check_launch("hgnn_lg_side_fwd");
}

// fill the derived fields of one backward part; returns false if it does not fit shared memory

static bool eng_plan_part(eng::BwdPart& p, int Fg, bool vec4, bool& vout4, bool is_self, bool wide) {
    p.nT = p.ops.n * Fg;
    p.P = p.nT * p.Fx;
    vout4 = vec4 && eng_part_vout4(p.X, p.gX, p.Fx);
    const int slots = wide ? (p.nT / 4) * (p.Fx / 4) : p.nT * (vout4 ? p.Fx / 4 : p.Fx);
    p.NG = slots >= ENG_THREADS ? 1 : ENG_THREADS / slots;
    const int tw = is_self ? p.nT + Fg : p.nT;
    p.Tp = eng_pad(tw, vec4 ? 4 : 1);
    p.Xp = vout4 ? eng_pad(p.Fx, 4) : (p.Fx | 1);
    const int rows_per_pass = ENG_THREADS / (vout4 ? p.Fx / 4 : p.Fx);
    int TR = wide ? 4 * rows_per_pass : min(rows_per_pass, max(32, ENG_THREADS / max(1, Fg / (vec4 ? 4 : 1))));
    const int PD = is_self ? p.P + Fg : p.P;
    auto smem_for = [&](int tr) {
        return ((size_t)((p.nT * p.Fx + 3) & ~3) + 4 * (size_t)((p.Fx + 3) & ~3) + (size_t)tr * p.Tp +
                (size_t)((tr * p.Xp + 3) & ~3) + (size_t)p.NG * PD) * sizeof(float);
    };
    while (TR > 1 && smem_for(TR) > ENG_MAX_SMEM) TR >>= 1;
    p.TR = TR;
    p.smem = smem_for(TR);
    p.tiles = ceil_div(p.R, TR);
    return p.smem <= ENG_MAX_SMEM;
}

// Width-only eligibility of one side on the engine kernels (forward AND backward): the same shared-memory
// arithmetic as hgnn_lg_side_fwd / hgnn_lg_side_bwd, assuming 16-byte aligned tensors (every torch allocation).
// engine.supported() asks this per side, so that a model whose weight block does not fit (LGNN order 1 from
// h ~ 48: Cin x Fout = 10h x 2h floats) runs on the per-layer kernels instead of failing inside a step.
extern "C" int hgnn_lg_side_fits(int n_ops, int Fs, int Fc, int Fout) {
    if (n_ops < 1 || n_ops > HGNN_MAX_OPS || Fs < 1 || Fs > 128 || Fc < 0 || Fc > 128 || Fout < 1 || Fout > 128) return 0;
    int TRw = 0;
    size_t smw = 0;
    // ---- forward
    bool fwd_ok = eng_wide_fwd_fits(n_ops, Fs, Fc, Fout, &TRw, &smw);
    if (!fwd_ok) {
        const bool vec4 = (Fs % 4 == 0) && (Fc % 4 == 0);
        const int Cin = n_ops * Fs + 2 * Fc;
        const size_t fixed = ((size_t)Cin * Fout + ((Fout + 3) & ~3) + 2 * ((Fs + 3) & ~3) + 2 * ((Fc + 3) & ~3)) * sizeof(float);
        fwd_ok = fixed + (size_t)eng_pad(Cin, vec4 ? 4 : 1) * sizeof(float) <= ENG_MAX_SMEM;
    }
    if (!fwd_ok) return 0;
    // ---- backward: both parts on the tensor-core tiles, or both on the generic tiles
    if (eng_wide_bwd_width(Fout) && eng_wide_part_fits(n_ops, Fout, Fs, true, &TRw, &smw) &&
        (Fc == 0 || eng_wide_part_fits(2, Fout, Fc, false, &TRw, &smw)))
        return 1;
    const bool vec4 = Fout % 4 == 0;
    const bool wide = vec4 && Fout >= 32 && Fs >= 16 && Fs % 4 == 0 && (Fc == 0 || (Fc >= 16 && Fc % 4 == 0));
    for (int part = 0; part < (Fc > 0 ? 2 : 1); ++part) {
        eng::BwdPart p;
        p.ops.n = part == 0 ? n_ops : 2;
        p.Fx = part == 0 ? Fs : Fc;
        p.R = 1;
        p.X = nullptr; p.gX = nullptr;      // null pointers count as aligned
        bool v4 = false;
        if (!eng_plan_part(p, Fout, vec4, v4, part == 0, wide)) return 0;
    }
    return 1;
}

// CTAs of a backward launch that work on the self rows.  Every thread walks WHOLE rows one after the other, so
// what matters is the integer number of rows of the slowest thread of each part: the split minimises
// max(ceil(rows_self / threads_self) * row cost_self, ceil(rows_cross / threads_cross) * row cost_cross); ties go to the
// split closest to the cost-proportional one.  (A proportional split left e.g. 1.3 rows per thread on the heavy
// part - a third of its threads did two rows and set the kernel time: profiles/logs/cta_times_*.log.)
static int eng_split_ctas_search(int grid, long long R_self, long long R_cross, double cost_s, double cost_c);
// The search walks every split of the grid (~600 candidates, two 64-bit divisions each: ~5 us of host time per
// backward launch, 0.15 ms per step); the 36 middle sides of a step ask the same two questions, so the last few
// answers are kept (per thread: no locking).
static int eng_split_ctas(int grid, long long R_self, long long R_cross, double cost_s, double cost_c) {
    if (R_cross <= 0) return grid;
    if (grid < 2) return 1;
    struct Memo { int grid; long long rs, rc; double cs, cc; int best; };
    static thread_local Memo memo[8];
    static thread_local int n_memo = 0, next = 0;
    for (int i = 0; i < n_memo; ++i) {
        const Memo& m = memo[i];
        if (m.grid == grid && m.rs == R_self && m.rc == R_cross && m.cs == cost_s && m.cc == cost_c) return m.best;
    }
    const int best = eng_split_ctas_search(grid, R_self, R_cross, cost_s, cost_c);
    Memo& m = memo[next];
    m.grid = grid; m.rs = R_self; m.rc = R_cross; m.cs = cost_s; m.cc = cost_c; m.best = best;
    next = (next + 1) & 7;
    if (n_memo < 8) ++n_memo;
    return best;
}
static int eng_split_ctas_search(int grid, long long R_self, long long R_cross, double cost_s, double cost_c) {
    const double row_s = cost_s / (double)R_self, row_c = cost_c / (double)R_cross;
    const double prop = grid * cost_s / (cost_s + cost_c);
    int best = 1;
    double best_t = 1e300, best_d = 1e300;
    for (int cs = 1; cs <= grid - 1; ++cs) {
        const long long ts = (long long)cs * R4_THREADS, tc = (long long)(grid - cs) * R4_THREADS;
        const double t_s = (double)((R_self + ts - 1) / ts) * row_s, t_c = (double)((R_cross + tc - 1) / tc) * row_c;
        const double t = t_s > t_c ? t_s : t_c;
        const double dist = cs > prop ? cs - prop : prop - cs;
        if (t < best_t - 1e-12 || (t < best_t + 1e-12 && dist < best_d)) { best_t = t; best_d = dist; best = cs; }
    }
    return best;
}

static bool eng_no_range_ctas() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HGNN_B200_NO_RANGE_CTAS"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// scratch of the dedicated range-sum CTAs: rng_n ready flags (ints, 16-byte padded) + 4 floats per range
extern "C" long long hgnn_lg_rng_scratch_bytes(int rng_n) {
    if (rng_n <= 0) return 0;
    return (((long long)rng_n * 4 + 15) & ~15ll) + (long long)rng_n * 16;
}

static bool eng_try_bwd_row4(const hgnn_side_bwd_t* d, hgnn_stream_t stream) {
    if (eng_row4_disabled()) return false;
    if (d->Fg != 4 || d->R_self <= 0 || d->Fs != 4 || !eng_row4_ops(d->ops_T, d->n_ops)) return false;
    if (d->R_cross > 0 && d->Fc != 4) return false;
    if (!eng_aligned16(d->gY) || (d->Z && !eng_aligned16(d->Z)) || !eng_aligned16(d->Xs) ||
        (d->gXs && !eng_aligned16(d->gXs)) || (d->R_cross > 0 && (!eng_aligned16(d->Xc) || (d->gXc && !eng_aligned16(d->gXc)))))
        return false;
    for (int i = 3; i < d->n_ops; ++i) if (d->ops_T[i].rng_rowptr) return false;   // ranges only on the first CSR op
    eng::Bwd4Args a;
    a.gY = d->gY; a.Z = d->Z; a.relu_from = d->relu_from; a.Rg = d->Rg; a.inv_Rg = d->Rg > 0 ? 1.0 / (double)d->Rg : 0.0; a.has_bn = d->acc_b != nullptr;
    a.acc_f = d->acc_f; a.acc_b = d->acc_b; a.bn_w = d->bn_weight;
    a.Wa = d->Wa; a.Ha = d->Ha; a.Wb = d->Wb; a.Hb = d->Hb; a.Cin = d->Cin;
    a.dW_bins = d->dW_bins; a.db_bins = d->db_bins;
    a.R_self = d->R_self; a.n_csr = d->n_ops - 2; a.diag = d->ops_T[1].diag;
    for (int i = 0; i < 2; ++i) {
        const bool on = i < a.n_csr;
        a.rowptr[i] = on ? d->ops_T[2 + i].rowptr : nullptr;
        a.col[i] = on ? d->ops_T[2 + i].col : nullptr;
        a.val[i] = on ? d->ops_T[2 + i].val : nullptr;
    }
    const hgnn_op_t& o2 = d->ops_T[2];
    a.rng_rowptr = o2.rng_rowptr; a.rng_id = o2.rng_id; a.rng_val = o2.rng_val; a.rng_lo = o2.rng_lo; a.rng_hi = o2.rng_hi;
    a.Xs = d->Xs; a.bn_s = to_bnref(&d->bn_self); a.gXs = d->gXs; a.acc_self = d->accumulate_self; a.acc_b_self = d->acc_b_self;
    a.R_cross = d->R_cross;
    a.pt_rowptr = d->pt_rowptr; a.pt_col = d->pt_col; a.pt_pm = d->pt_pm; a.pt_pd = d->pt_pd;
    a.Xc = d->Xc; a.bn_c = to_bnref(&d->bn_cross); a.gXc = d->gXc; a.acc_cross = d->accumulate_cross;
    a.acc_b_cross = d->acc_b_cross; a.col0_cross = d->n_ops * 4;
    a.roww_s = d->roww_self; a.roww_c = d->roww_cross; a.rowmap_s = d->rowmap_self; a.rowmap_c = d->rowmap_cross;
    // dedicated range-sum CTAs when the caller provided the (zeroed) scratch: [flags: rng_n ints | sums: 4 floats each]
    a.rng_n = 0; a.range_ctas = 0; a.rng_sum_g = nullptr; a.rng_flag_g = nullptr;
    if (d->rng_scratch && o2.rng_rowptr && o2.rng_n > 0 && !eng_no_range_ctas()) {
        a.rng_n = o2.rng_n;
        a.range_ctas = o2.rng_n < 32 ? o2.rng_n : 32;
        a.rng_flag_g = static_cast<int*>(d->rng_scratch);
        a.rng_sum_g = reinterpret_cast<float*>(static_cast<char*>(d->rng_scratch) + (((size_t)o2.rng_n * 4 + 15) & ~(size_t)15));
    }
    static int ablate = -1;
    if (ablate < 0) { const char* e = getenv("HGNN_B200_ABLATE"); ablate = e ? atoi(e) : 0; }
    a.ablate = ablate & ~16;
    a.trace_slot = -1;
    cudaStream_t s = to_stream(stream);
    const long long rows = (long long)d->R_self + (d->R_cross > 0 ? d->R_cross : 0);
    // The CTAs of one launch are split between the two parts in proportion to their estimated cost, not
    // their row counts: a row costs about one unit plus a third of a unit per gathered entry (two row
    // loads each).  At the C2 workload the edge side differentiates 160 k line-graph rows with 0.5
    // entries each against 32 k node rows with 8 entries each - a split by rows left the node rows
    // with 17 % of the CTAs and 3x the work per thread.
    const double avg_s = (double)d->ops_T[2].nnz / d->R_self;
    const double avg_c = d->R_cross > 0 ? (double)d->pt_nnz / d->R_cross : 0.0;
    static double w_self = -1.0, w_cross = -1.0;
    if (w_self < 0.0) {
        w_self = 0.6; w_cross = 0.3;      // per-row cost ratios measured with profiles/cta_times.py
        const char* e = getenv("HGNN_B200_BWD_ENTRY_COST");      // "w_self,w_cross" (tuning aid)
        if (e) { w_self = atof(e); const char* c = strchr(e, ','); w_cross = c ? atof(c + 1) : w_self; }
    }
    const double act_s = (double)d->R_self, act_c = (double)d->R_cross;
    const double cost_s = act_s * 1.0 + w_self * (double)d->ops_T[2].nnz;
    const double cost_c = d->R_cross > 0 ? act_c * 1.0 + w_cross * (double)d->pt_nnz : 0.0;
    static int debug_split = -1;
    if (debug_split < 0) debug_split = getenv("HGNN_B200_DEBUG_SPLIT") ? 1 : 0;
    bool big_s = false, big_c = false;      // measured: the small batches win in the backward (register pressure)
    static int bforce = -2;
    if (bforce == -2) { const char* e = getenv("HGNN_B200_BWD_BATCH"); bforce = e ? atoi(e) : -1; }  // 0: (2,4), 1: (8,4), 2: (2,8), 3: (4,4)
    bool mid_s = false;
    if (bforce >= 0) { big_s = bforce == 1; big_c = bforce == 2; mid_s = bforce == 3; }
#define R4_BWD(NCSR, DW, GB, CB)                                                                          \
    {                                                                                                     \
        const int cap = eng_resident_impl((const void*)eng::bwd_row4_kernel<NCSR, DW, GB, CB>, 0, R4_THREADS); \
        int grid = (int)min((long long)cap, (rows + R4_THREADS - 1) / R4_THREADS);                        \
        if (d->R_cross > 0 && grid < 2) grid = 2;                                                         \
        int row_grid = grid;                            /* the range CTAs come out of the resident budget */ \
        if (a.range_ctas > 0 && grid + a.range_ctas > cap) row_grid = max(2, cap - a.range_ctas);         \
        a.ctas_self = eng_split_ctas(row_grid, (long long)act_s, d->R_cross > 0 ? (long long)act_c : 0, cost_s, cost_c); \
        if (debug_split)                                                                                  \
            fprintf(stderr, "bwd_row4 split: cap %d grid %d row_grid %d range_ctas %d R_self %d R_cross %d avg_s %.3f avg_c %.3f -> ctas_self %d\n", \
                    cap, grid, row_grid, a.range_ctas, d->R_self, d->R_cross, avg_s, avg_c, a.ctas_self); \
        eng_launch(eng::bwd_row4_kernel<NCSR, DW, GB, CB>, row_grid + a.range_ctas, R4_THREADS, 0, s, a); \
    }
    // low-register variant: one CSR operator without a run-length part (A^T, or AL^T of the collapsed line graph)
    // Opt-in (HGNN_B200_BWD_V2=1): measured on C2 it is 0.4 us faster on the node side (11.9 vs 12.3 us) and 5 us SLOWER
    // on the edge side (18.7 vs 13.7 us) - the per-row tile traffic (24 STS + 32 LDS.128 + 64 FMA) outweighs the higher
    // occupancy on the light line-graph rows: 0.824 vs 0.754 ms per step (profiles/logs/bench_r2i_*.log).
    static int v2 = -1;
    if (v2 < 0) { const char* e = getenv("HGNN_B200_BWD_V2"); v2 = (e && e[0] == '1') ? 1 : 0; }
    if (v2 && a.n_csr == 1 && !a.rng_rowptr && !d->skip_dw) {
        const bool big_cross = avg_c > 5.0;
#define R4C_BWD(GB, CB)                                                                                   \
        {                                                                                                 \
            const int cap = eng_resident_impl((const void*)eng::bwd_row4c_kernel<GB, CB>, 0, R4_THREADS); \
            int grid = (int)min((long long)cap, (rows + R4_THREADS - 1) / R4_THREADS);                    \
            if (d->R_cross > 0 && grid < 2) grid = 2;                                                     \
            a.ctas_self = eng_split_ctas(grid, (long long)act_s, d->R_cross > 0 ? (long long)act_c : 0, cost_s, cost_c); \
            if (debug_split)                                                                              \
                fprintf(stderr, "bwd_row4c split: cap %d grid %d R_self %d R_cross %d avg_s %.3f avg_c %.3f -> ctas_self %d\n", \
                        cap, grid, d->R_self, d->R_cross, avg_s, avg_c, a.ctas_self);                     \
            eng_launch(eng::bwd_row4c_kernel<GB, CB>, grid, R4_THREADS, 0, s, a);                         \
        }
        if (big_cross) R4C_BWD(4, 8) else R4C_BWD(4, 4)
#undef R4C_BWD
        return true;
    }
    // pipelined variant (engine_row4p.cuh): one CSR operator, no run-length part, weight gradients in the same launch.
    // HGNN_B200_BWD_P=0 switches back to bwd_row4_kernel.
    static int pvar = -1;
    if (pvar < 0) { const char* e = getenv("HGNN_B200_BWD_P"); pvar = (e && e[0] == '0') ? 0 : 1; }
    if (pvar && a.n_csr == 1 && !a.rng_rowptr && a.range_ctas == 0 && !d->skip_dw && !(a.ablate & 7)) {
#define R4P_BWD_(GB, CB, RMW)                                                                                  \
        {                                                                                                 \
            const int cap = eng_resident_impl((const void*)eng::bwd_row4p_kernel<GB, CB, RMW>, 0, R4_THREADS); \
            int grid = (int)min((long long)cap, (rows + R4_THREADS - 1) / R4_THREADS);                    \
            if (d->R_cross > 0 && grid < 2) grid = 2;                                                     \
            a.ctas_self = eng_split_ctas(grid, (long long)act_s, d->R_cross > 0 ? (long long)act_c : 0, cost_s, cost_c); \
            if (d->R_cross <= 0) a.ctas_self = grid;                                                      \
            if (debug_split)                                                                              \
                fprintf(stderr, "bwd_row4p split: cap %d grid %d R_self %d R_cross %d avg_s %.3f avg_c %.3f -> ctas_self %d\n", \
                        cap, grid, d->R_self, d->R_cross, avg_s, avg_c, a.ctas_self);                     \
            a.trace_slot = eng_trace_slot(ablate);                                                        \
            eng_launch(eng::bwd_row4p_kernel<GB, CB, RMW>, grid, R4_THREADS, 0, s, a);                         \
        }
        // batch sizes: 2 entries for the transposed CSR operator (more spills under the 128-register bound); the incidence
        // pattern takes 2 when its rows are line-graph nodes (two end points each: a listed cross part, whatever the average
        // over all rows says), else 4.  HGNN_B200_BWD_BATCH: 3 = (4, 4),
        // 2 = (2, 8) (both spill; measurement only), 5 = (2, 2), 0 = (2, 4).
        static int rmw = -1;
        if (rmw < 0) { const char* e = getenv("HGNN_B200_BWD_RMW"); rmw = (e && e[0] == '1') ? 1 : 0; }   // measured equal on C2 (0.634 vs 0.630 ms per step); the reduction form spills less
#define R4P_BWD(GB, CB) { if (rmw) R4P_BWD_(GB, CB, true) else R4P_BWD_(GB, CB, false) }
        static int edge44 = -1;
        if (edge44 < 0) { const char* e = getenv("HGNN_B200_BWD_EDGE44"); edge44 = (e && e[0] == '1') ? 1 : 0; }
        if (edge44 && d->rowmap_self && bforce < 0) mid_s = true;      // experiment: (4, 4) on the collapsed edge side only
        if (mid_s) R4P_BWD(4, 4) else if (big_c) R4P_BWD(2, 8) else if (bforce == 5 || (bforce < 0 && (avg_c <= 2.5 || d->rowmap_cross))) R4P_BWD(2, 2) else R4P_BWD(2, 4)
#undef R4P_BWD
#undef R4P_BWD_
        return true;
    }
#define R4_BWD_B(NCSR, DW)                                                                                \
    if (big_s) R4_BWD(NCSR, DW, 8, 4) else if (big_c) R4_BWD(NCSR, DW, 2, 8) else if (mid_s) R4_BWD(NCSR, DW, 4, 4) else R4_BWD(NCSR, DW, 2, 4)
    if (d->skip_dw) { if (a.n_csr == 1) { R4_BWD_B(1, false) } else { R4_BWD_B(2, false) } }
    else { if (a.n_csr == 1) { R4_BWD_B(1, true) } else { R4_BWD_B(2, true) } }
#undef R4_BWD_B
#undef R4_BWD
    return true;
}

// thread-per-row backward for the odd small widths of layer 0 and the readout (engine_rowg.cuh)
static bool eng_try_bwd_rowg(const hgnn_side_bwd_t* d, hgnn_stream_t stream) {
    if (eng_row4_disabled() || d->skip_dw) return false;
    const int Fg = d->Fg, Fs = d->Fs, Fc = d->R_cross > 0 ? d->Fc : 0;
    if (d->R_self <= 0 || !d->ops_T || !eng_row4_ops(d->ops_T, d->n_ops)) return false;
    for (int i = 3; i < d->n_ops; ++i) if (d->ops_T[i].rng_rowptr) return false;   // ranges only on the first CSR op
    const bool has_bn = d->acc_b != nullptr;
    if (Fg < 4 && (has_bn || d->relu_from < Fg)) return false;
    if ((d->bn_self.acc || d->bn_self.affine) && Fs != 4) return false;
    if (d->R_cross > 0 && (d->bn_cross.acc || d->bn_cross.affine) && Fc != 4) return false;
    if (Fg == 4 && (!eng_aligned16(d->gY) || (d->Z && !eng_aligned16(d->Z)))) return false;
    if ((Fs == 4 && !eng_aligned16(d->Xs)) || (Fc == 4 && !eng_aligned16(d->Xc))) return false;
    eng::Bwd4Args a;
    a.gY = d->gY; a.Z = d->Z; a.relu_from = d->relu_from; a.Rg = d->Rg; a.inv_Rg = d->Rg > 0 ? 1.0 / (double)d->Rg : 0.0; a.has_bn = has_bn;
    a.acc_f = d->acc_f; a.acc_b = d->acc_b; a.bn_w = d->bn_weight;
    a.Wa = d->Wa; a.Ha = d->Ha; a.Wb = d->Wb; a.Hb = d->Hb; a.Cin = d->Cin;
    a.dW_bins = d->dW_bins; a.db_bins = d->db_bins;
    a.R_self = d->R_self; a.n_csr = d->n_ops - 2; a.diag = d->ops_T[1].diag;
    for (int i = 0; i < 2; ++i) {
        const bool on = i < a.n_csr;
        a.rowptr[i] = on ? d->ops_T[2 + i].rowptr : nullptr;
        a.col[i] = on ? d->ops_T[2 + i].col : nullptr;
        a.val[i] = on ? d->ops_T[2 + i].val : nullptr;
    }
    const hgnn_op_t& o2 = d->ops_T[2];
    a.rng_rowptr = o2.rng_rowptr; a.rng_id = o2.rng_id; a.rng_val = o2.rng_val; a.rng_lo = o2.rng_lo; a.rng_hi = o2.rng_hi;
    a.Xs = d->Xs; a.bn_s = to_bnref(&d->bn_self); a.gXs = d->gXs; a.acc_self = d->accumulate_self; a.acc_b_self = d->acc_b_self;
    a.R_cross = d->R_cross > 0 ? d->R_cross : 0;
    a.pt_rowptr = d->pt_rowptr; a.pt_col = d->pt_col; a.pt_pm = d->pt_pm; a.pt_pd = d->pt_pd;
    a.Xc = d->Xc; a.bn_c = to_bnref(&d->bn_cross); a.gXc = d->gXc; a.acc_cross = d->accumulate_cross;
    a.acc_b_cross = d->acc_b_cross; a.col0_cross = d->n_ops * Fs; a.ablate = 0; a.trace_slot = -1;
    a.roww_s = d->roww_self; a.roww_c = d->roww_cross; a.rowmap_s = d->rowmap_self; a.rowmap_c = d->rowmap_cross;
    a.rng_n = 0; a.range_ctas = 0; a.rng_sum_g = nullptr; a.rng_flag_g = nullptr;
    cudaStream_t s = to_stream(stream);
    const long long rows = (long long)d->R_self + a.R_cross;
    const double avg_s = (double)d->ops_T[2].nnz / d->R_self;
    const double avg_c = a.R_cross > 0 ? (double)d->pt_nnz / a.R_cross : 0.0;
    const double cost_s = (double)d->R_self * (1.0 + 0.6 * avg_s);
    const double cost_c = a.R_cross > 0 ? (double)a.R_cross * (1.0 + 0.3 * avg_c) : 0.0;
    bool done = false;
#define RG_BWD(NCSR, FG, FS, FC)                                                                           \
    if (!done && a.n_csr == NCSR && Fg == FG && Fs == FS && Fc == FC) {                                    \
        const int cap = eng_resident_impl((const void*)eng::bwd_rowg_kernel<NCSR, FG, FS, FC>, 0, R4_THREADS); \
        int grid = (int)min((long long)cap, (rows + R4_THREADS - 1) / R4_THREADS);                         \
        if (a.R_cross > 0 && grid < 2) grid = 2;                                                           \
        a.ctas_self = eng_split_ctas(grid, d->R_self, a.R_cross, cost_s, cost_c);                          \
        eng_launch(eng::bwd_rowg_kernel<NCSR, FG, FS, FC>, grid, R4_THREADS, 0, s, a);                     \
        done = true;                                                                                       \
    }
    // LGNN: layer 0 node side, layer 0 edge side, readout (dim_output 2 / 1); power GNN: layer 0, readout
    RG_BWD(1, 4, 5, 1) RG_BWD(1, 4, 1, 4) RG_BWD(1, 2, 4, 4) RG_BWD(1, 1, 4, 4) RG_BWD(1, 4, 5, 0) RG_BWD(1, 2, 4, 0) RG_BWD(1, 1, 4, 0)
    // J = 2: the layer-0 node side would need 80 dW accumulators per thread (> 64): generic tile kernels
    RG_BWD(2, 4, 1, 4) RG_BWD(2, 2, 4, 4) RG_BWD(2, 1, 4, 4) RG_BWD(2, 2, 4, 0) RG_BWD(2, 1, 4, 0)
#undef RG_BWD
    return done;
}

// profiling aid: copies (start ns, end ns, is_self) of the first n CTAs of the last width-4 backward launch
extern "C" int hgnn_debug_cta_times(unsigned long long* out, int n) {
    HGNN_REQUIRE(out && n > 0 && n <= 2048, "bad argument");
    cudaError_t e = cudaMemcpyFromSymbol(out, eng::g_cta_times, (size_t)n * 3 * sizeof(unsigned long long));
    if (e != cudaSuccess) { hgnn_set_error("hgnn_debug_cta_times: %s", cudaGetErrorString(e)); return HGNN_ERR_CUDA; }
    return HGNN_OK;
}

// (end of the row loop ns, end of the range phase ns, coefficient vectors ready ns) of the row CTAs of the same launch
extern "C" int hgnn_debug_cta_phases(unsigned long long* out, int n) {
    HGNN_REQUIRE(out && n > 0 && n <= 2048, "bad argument");
    cudaError_t e = cudaMemcpyFromSymbol(out, eng::g_cta_phase, (size_t)n * 3 * sizeof(unsigned long long));
    if (e != cudaSuccess) { hgnn_set_error("hgnn_debug_cta_phases: %s", cudaGetErrorString(e)); return HGNN_ERR_CUDA; }
    return HGNN_OK;
}

// Launch timeline (HGNN_B200_ABLATE bit 16).  out != NULL: copies the first n slots (4 values each: min CTA start, min /
// max "wait passed", max CTA end; ns) to the HOST array.  reset = 1: slots back to (max, max, 0, 0) and the next traced
// launch takes slot 0 again; reset = 2: slot values only (before re-running a captured graph whose launches keep their slots).
extern "C" int hgnn_debug_ktrace(unsigned long long* out, int n, int reset) {
    HGNN_REQUIRE(n >= 0 && n <= KTRACE_SLOTS, "bad argument");
    if (out && n > 0) {
        cudaError_t e = cudaMemcpyFromSymbol(out, eng::g_ktrace, (size_t)n * 4 * sizeof(unsigned long long));
        if (e != cudaSuccess) { hgnn_set_error("hgnn_debug_ktrace: %s", cudaGetErrorString(e)); return HGNN_ERR_CUDA; }
    }
    if (reset) {
        static unsigned long long init[KTRACE_SLOTS * 4];
        for (int i = 0; i < KTRACE_SLOTS; ++i) { init[i * 4] = init[i * 4 + 1] = ~0ull; init[i * 4 + 2] = init[i * 4 + 3] = 0ull; }
        cudaError_t e = cudaMemcpyToSymbol(eng::g_ktrace, init, sizeof(init));
        if (e != cudaSuccess) { hgnn_set_error("hgnn_debug_ktrace: %s", cudaGetErrorString(e)); return HGNN_ERR_CUDA; }
        if (reset == 1) g_ktrace_next = 0;
    }
    return HGNN_OK;
}

extern "C" int hgnn_lg_side_dw(const float* gY, const float* Z, int R, int relu_from, const double* acc_f,
                               const double* acc_b, const float* bn_weight, const float* X1, int Cin,
                               double* dW_bins, double* db_bins, hgnn_stream_t stream) {
    HGNN_REQUIRE(gY && X1 && dW_bins && R >= 0 && Cin % 4 == 0 && Cin >= 8 && Cin <= 24, "bad argument");
    HGNN_REQUIRE(!acc_b || (acc_f && bn_weight && Z), "batch-norm backward needs acc_f, bn_weight and Z");
    HGNN_REQUIRE(relu_from >= 4 || Z, "ReLU backward needs Z");
    HGNN_REQUIRE(eng_aligned16(gY) && eng_aligned16(X1) && (!Z || eng_aligned16(Z)), "misaligned rows");
    if (R == 0) return HGNN_OK;
    eng::Dw4Args a;
    a.gY = gY; a.Z = Z; a.relu_from = relu_from; a.R = R; a.has_bn = acc_b != nullptr;
    a.acc_f = acc_f; a.acc_b = acc_b; a.bn_w = bn_weight; a.X1 = X1; a.Cin = Cin;
    a.dW_bins = dW_bins; a.db_bins = db_bins;
    cudaStream_t s = to_stream(stream);
    const int grid = min(ceil_div(R, R4_THREADS), HGNN_SM_COUNT * 4);
    switch (Cin / 4) {
        case 2: eng::dw_row4_kernel<2><<<grid, R4_THREADS, 0, s>>>(a); break;
        case 3: eng::dw_row4_kernel<3><<<grid, R4_THREADS, 0, s>>>(a); break;
        case 4: eng::dw_row4_kernel<4><<<grid, R4_THREADS, 0, s>>>(a); break;
        case 5: eng::dw_row4_kernel<5><<<grid, R4_THREADS, 0, s>>>(a); break;
        default: eng::dw_row4_kernel<6><<<grid, R4_THREADS, 0, s>>>(a); break;
    }
    return hgnn_check_launch("hgnn_lg_side_dw");
}

extern "C" int hgnn_lg_side_bwd(const hgnn_side_bwd_t* d, hgnn_stream_t stream) {
    HGNN_REQUIRE(d && d->gY && d->Fg >= 1 && d->Fg <= 128, "bad argument");
    HGNN_REQUIRE(!d->skip_dw || (d->Fg == 4 && d->Fs == 4), "skip_dw is only valid on the width-4 fast path");
    HGNN_REQUIRE(d->Ha >= 0 && d->Hb >= 0 && d->Ha + d->Hb == d->Fg, "Ha + Hb must equal the width of gY");
    HGNN_REQUIRE((d->Ha == 0 || d->Wa) && (d->Hb == 0 || d->Wb), "null weights");
    HGNN_REQUIRE(!d->acc_b || (d->acc_f && d->bn_weight && d->Z), "batch-norm backward needs acc_f, bn_weight and Z");
    HGNN_REQUIRE(d->relu_from >= d->Fg || d->Z, "ReLU backward needs Z");
    if (eng_try_bwd_row4(d, stream)) return hgnn_check_launch("hgnn_lg_side_bwd(row4)");
    if (eng_try_bwd_rowg(d, stream)) return hgnn_check_launch("hgnn_lg_side_bwd(rowg)");
    HGNN_REQUIRE(!eng_recording(), "recording: this side does not run on the thread-per-row kernels");
    HGNN_REQUIRE(!d->skip_dw, "skip_dw needs the width-4 fast path (check hgnn_lg_row4_eligible)");
    HGNN_REQUIRE(!d->roww_self && !d->roww_cross && !d->rowmap_self && !d->rowmap_cross, "row weights need the thread-per-row kernels (widths of the h = 2 feature maps)");
    eng::BwdArgs a;
    a.gY = d->gY; a.Z = d->Z; a.Fg = d->Fg; a.relu_from = d->relu_from; a.Rg = d->Rg;
    a.acc_f = d->acc_f; a.acc_b = d->acc_b; a.bn_w = d->bn_weight;
    a.Wa = d->Wa; a.Ha = d->Ha; a.Wb = d->Wb; a.Hb = d->Hb; a.Cin = d->Cin;
    a.dW_bins = d->dW_bins; a.db_bins = d->db_bins;
    const bool vec4 = (d->Fg % 4 == 0) && eng_aligned16(d->gY) && (!d->Z || eng_aligned16(d->Z));
    // wide states: register-tiled phases (both parts must take the float4 paths)
    const bool wide = vec4 && d->Fg >= 32 &&
        (d->R_self <= 0 || (d->Fs >= 16 && d->Xs && eng_part_vout4(d->Xs, d->gXs, d->Fs))) &&
        (d->R_cross <= 0 || (d->Fc >= 16 && d->Xc && eng_part_vout4(d->Xc, d->gXc, d->Fc)));
    // tensor-core tile kernel (engine_wide.cuh) when both parts qualify
    bool mma = vec4 && eng_wide_bwd_width(d->Fg);
    // self part
    a.self.R = d->R_self;
    bool vs4 = false, vc4 = false;
    if (d->R_self > 0) {
        HGNN_REQUIRE(d->ops_T && make_oplist(d->ops_T, d->n_ops, &a.self.ops) == 0 && d->n_ops >= 1, "bad operator list");
        HGNN_REQUIRE(d->Xs && d->Fs >= 1 && d->Fs <= 128, "bad self input");
        a.self.X = d->Xs; a.self.Fx = d->Fs; a.self.bn = to_bnref(&d->bn_self);
        a.self.gX = d->gXs; a.self.accumulate = d->accumulate_self; a.self.acc_b = d->acc_b_self;
        a.self.col0 = 0;
        if (mma) mma = eng_plan_part_wide(a.self, d->Fg, true);
        if (!mma && !eng_plan_part(a.self, d->Fg, vec4, vs4, true, wide)) {
            hgnn_set_error("hgnn_lg_side_bwd: self block %d x %d does not fit shared memory", a.self.nT, d->Fs);
            return HGNN_ERR_ARG;
        }
    } else {
        a.self.tiles = 0; a.self.smem = 0; a.self.ops.n = 0;
    }
    a.cross.R = d->R_cross;
    if (d->R_cross > 0) {
        HGNN_REQUIRE(d->pt_rowptr && d->pt_col && d->pt_pm && d->pt_pd && d->Xc && d->Fc >= 1 && d->Fc <= 128, "bad cross input");
        hgnn_op_t cops[2] = {};
        for (int i = 0; i < 2; ++i) {
            cops[i].kind = HGNN_OP_CSR; cops[i].diag = nullptr;
            cops[i].rowptr = d->pt_rowptr; cops[i].col = d->pt_col;
            cops[i].val = i == 0 ? d->pt_pm : d->pt_pd;
        }
        make_oplist(cops, 2, &a.cross.ops);
        a.cross.X = d->Xc; a.cross.Fx = d->Fc; a.cross.bn = to_bnref(&d->bn_cross);
        a.cross.gX = d->gXc; a.cross.accumulate = d->accumulate_cross; a.cross.acc_b = d->acc_b_cross;
        a.cross.col0 = d->n_ops * d->Fs;
        if (mma && !eng_plan_part_wide(a.cross, d->Fg, false)) {
            // the cross part does not qualify: plan both parts for the generic kernel
            mma = false;
            if (d->R_self > 0 && !eng_plan_part(a.self, d->Fg, vec4, vs4, true, wide)) {
                hgnn_set_error("hgnn_lg_side_bwd: self block %d x %d does not fit shared memory", a.self.nT, d->Fs);
                return HGNN_ERR_ARG;
            }
        }
        if (!mma && !eng_plan_part(a.cross, d->Fg, vec4, vc4, false, wide)) {
            hgnn_set_error("hgnn_lg_side_bwd: cross block %d x %d does not fit shared memory", a.cross.nT, d->Fc);
            return HGNN_ERR_ARG;
        }
    } else {
        a.cross.tiles = 0; a.cross.smem = 0; a.cross.ops.n = 0;
    }
    const int total_tiles = a.self.tiles + a.cross.tiles;
    if (total_tiles == 0) return HGNN_OK;
    const size_t smem = a.self.smem > a.cross.smem ? a.self.smem : a.cross.smem;
    cudaStream_t s = to_stream(stream);
    // EXPERIMENTAL tcgen05 backward (engine_tc5.cuh): only with HGNN_B200_WIDE_TC5_BWD=1
    if (eng_tc5_bwd_enabled() && !eng_wide_disabled() && vec4 && d->Fg == T5B_F) {
        bool ok = true;
        if (a.self.R > 0) {
            ok = ok && a.self.Fx == T5B_F && a.self.ops.n <= 3 && eng_part_vout4(a.self.X, a.self.gX, a.self.Fx);
            for (int t = 0; ok && t < a.self.ops.n; ++t) ok = !a.self.ops.rng_rowptr[t];
        }
        if (a.cross.R > 0) ok = ok && a.cross.Fx == T5B_F && eng_part_vout4(a.cross.X, a.cross.gX, a.cross.Fx);
        if (ok) {
            int nT = 0;
            if (a.self.R > 0) { a.self.TR = 64; a.self.tiles = ceil_div(a.self.R, 64); nT = a.self.ops.n * T5B_F; }
            if (a.cross.R > 0) { a.cross.TR = 64; a.cross.tiles = ceil_div(a.cross.R, 64); if (2 * T5B_F > nT) nT = 2 * T5B_F; }
            const size_t sm5 = (size_t)eng::tc5_bwd_layout(nT).total * sizeof(float);
            if (sm5 <= WD_MAX_SMEM) {
                const int most = a.self.tiles > a.cross.tiles ? a.self.tiles : a.cross.tiles;
                const int grid = balanced_grid(most, eng_resident_impl(reinterpret_cast<const void*>(eng::bwd_tc5_kernel), sm5, WD_THREADS));
                eng::bwd_tc5_kernel<<<grid, WD_THREADS, sm5, s>>>(a);
                return hgnn_check_launch("hgnn_lg_side_bwd(tc5)");
            }
        }
    }
    if (mma) {
        // every CTA works on both parts (see bwd_wide_kernel)
        const int most = a.self.tiles > a.cross.tiles ? a.self.tiles : a.cross.tiles;
        const int grid = balanced_grid(most, eng_resident_impl(reinterpret_cast<const void*>(eng::bwd_wide_kernel), smem, WD_THREADS));
        eng::bwd_wide_kernel<<<grid, WD_THREADS, smem, s>>>(a);
        return hgnn_check_launch("hgnn_lg_side_bwd(wide)");
    }
#define ENG_LAUNCH(VEC, VS, VC, W)                                                                   \
    {                                                                                                \
        int grid = balanced_grid(total_tiles, eng_resident(eng::bwd_kernel<VEC, VS, VC, W>, smem));  \
        if (a.self.tiles > 0 && a.cross.tiles > 0 && grid < 2) grid = 2;                             \
        eng::bwd_kernel<VEC, VS, VC, W><<<grid, ENG_THREADS, smem, s>>>(a);                          \
    }
    if (wide) {
        ENG_LAUNCH(4, 4, 4, true)
    } else if (vec4) {
        if (vs4 && vc4) ENG_LAUNCH(4, 4, 4, false)
        else if (vs4) ENG_LAUNCH(4, 4, 1, false)
        else if (vc4) ENG_LAUNCH(4, 1, 4, false)
        else ENG_LAUNCH(4, 1, 1, false)
    } else {
        ENG_LAUNCH(1, 1, 1, false)
    }
#undef ENG_LAUNCH
    return hgnn_check_launch("hgnn_lg_side_bwd");
}

// ---------------------------------------------------------------------------------------------
// step-end reductions
// ---------------------------------------------------------------------------------------------
// out[i] = sum_{b < nb[i]} sum_{c < cnt[i]} arena[off[i] + b*stride[i] + c]
__global__ void bins_reduce_kernel(const double* __restrict__ arena, const long long* __restrict__ off,
                                   const int* __restrict__ nb, const int* __restrict__ stride,
                                   const int* __restrict__ cnt, int n, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* p = arena + off[i];
        const int s = stride[i], k = cnt[i], B = nb[i];
        double t = 0.0;
        for (int b = 0; b < B; ++b)
            for (int c = 0; c < k; ++c) t += p[(size_t)b * s + c];
        out[i] = (float)t;
    }
}

extern "C" int hgnn_bins_reduce(const double* arena, const long long* off, const int* nb, const int* stride,
                                const int* cnt, int n, float* out, hgnn_stream_t stream) {
    HGNN_REQUIRE(arena && off && nb && stride && cnt && out && n >= 0, "bad argument");
    if (n == 0) return HGNN_OK;
    bins_reduce_kernel<<<min(ceil_div(n, 128), HGNN_MAX_GRID), 128, 0, to_stream(stream)>>>(arena, off, nb, stride, cnt, n, out);
    return hgnn_check_launch("hgnn_bins_reduce");
}

// running statistics of every BN of the model in one launch: BN k has width F[k], accumulators at
// arena + acc_off[k] (binned (sum z, sum z^2)), n_rows[k] rows, and running [mean(F) | std(F)] at
// running + run_off[k]:  r = (1-momentum)*batch + momentum*r   (batch_normalization.py:37-38)
__global__ void bn_running_kernel(const double* __restrict__ arena, const long long* __restrict__ acc_off,
                                  const int* __restrict__ F, const int* __restrict__ n_rows,
                                  const int* __restrict__ rows_kind, int Rn, int Rm,
                                  const long long* __restrict__ run_off, int n_bn, float momentum,
                                  float* __restrict__ running) {
    const int k = blockIdx.x;
    if (k >= n_bn) return;
    const int Fk = F[k], nb = hgnn_ws_bins(2 * Fk);
    const double* acc = arena + acc_off[k];
    float* rm = running + run_off[k];
    float* rs = rm + Fk;
    for (int f = threadIdx.x; f < Fk; f += blockDim.x) {
        double a = 0.0, b = 0.0;
        for (int bin = 0; bin < nb; ++bin) {
            a += acc[(size_t)bin * 2 * Fk + f];
            b += acc[(size_t)bin * 2 * Fk + Fk + f];
        }
        const double n = (double)(n_rows ? n_rows[k] : (rows_kind[k] ? Rm : Rn));
        const double m = a / n;
        double var = b / n - m * m;
        if (var < 0.0) var = 0.0;
        const double sd = sqrt(var + ENG_BN_EPS);
        rm[f] = (1.f - momentum) * (float)m + momentum * rm[f];
        rs[f] = (1.f - momentum) * (float)sd + momentum * rs[f];
    }
}

extern "C" int hgnn_bn_running_update(const double* arena, const long long* acc_off, const int* F,
                                      const int* n_rows, const long long* run_off, int n_bn,
                                      float momentum, float* running, hgnn_stream_t stream) {
    HGNN_REQUIRE(arena && acc_off && F && n_rows && run_off && running && n_bn >= 0, "bad argument");
    if (n_bn == 0) return HGNN_OK;
    bn_running_kernel<<<n_bn, 64, 0, to_stream(stream)>>>(arena, acc_off, F, n_rows, nullptr, 0, 0, run_off, n_bn,
                                                          momentum, running);
    return hgnn_check_launch("hgnn_bn_running_update");
}

extern "C" int hgnn_bn_running_update_k(const double* arena, const long long* acc_off, const int* F,
                                        const int* rows_kind, int Rn, int Rm, const long long* run_off, int n_bn,
                                        float momentum, float* running, hgnn_stream_t stream) {
    HGNN_REQUIRE(arena && acc_off && F && rows_kind && run_off && running && n_bn >= 0, "bad argument");
    if (n_bn == 0) return HGNN_OK;
    bn_running_kernel<<<n_bn, 64, 0, to_stream(stream)>>>(arena, acc_off, F, nullptr, rows_kind, Rn, Rm, run_off, n_bn,
                                                          momentum, running);
    return hgnn_check_launch("hgnn_bn_running_update_k");
}

// readout backward prologue (layers_mnb.py:92,:386): G[r, o] = g[graph(r), o], and the bias gradient
// of the padded slots, sum_b pad_count[b] * g[b, o], goes straight into the dbias accumulators.
__global__ void readout_bwd_prep_kernel(const float* __restrict__ g, int bs, int F, const int* __restrict__ off,
                                        const float* __restrict__ pad_count, float* __restrict__ G,
                                        double* db_bins) {
    const int b = blockIdx.y;
    const int r0 = off[b];
    const long long n = (long long)(off[b + 1] - r0) * F;
    float* dst = G + (size_t)r0 * F;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = g[(size_t)b * F + (int)(i % F)];
    if (db_bins && pad_count && blockIdx.x == 0 && (int)threadIdx.x < F)
        atomicAdd(db_bins + threadIdx.x, (double)pad_count[b] * (double)g[(size_t)b * F + threadIdx.x]);
}

extern "C" int hgnn_readout_bwd_prep(const float* g, int bs, int F, const int* off, const float* pad_count,
                                     float* G, double* db_bins, hgnn_stream_t stream) {
    HGNN_REQUIRE(g && off && G && bs >= 0 && F >= 1 && F <= 256, "bad argument");
    if (bs == 0) return HGNN_OK;
    HGNN_REQUIRE(bs <= 65535, "bs > 65535");
    dim3 grid(16, bs);
    readout_bwd_prep_kernel<<<grid, 256, 0, to_stream(stream)>>>(g, bs, F, off, pad_count, G, db_bins);
    return hgnn_check_launch("hgnn_readout_bwd_prep");
}
