// side.cu -- the fused layer-side kernels: multi-operator gather -> concat -> two 1x1 convs ->
// ReLU -> batch-norm statistics (forward) and the transposed gather -> W^T -> dW (backward).
//
// Reference: layer_simple / layer_with_lg_{1,2,3} / layer_last(_lg).forward
// (models/layers/layers_mnb.py:52-69, 189-225, 256-290, 322-358, 88-95, 379-388), i.e.
// graph_oper + P_multi + torch.cat + Conv1d(k=1) x2 + ReLU + BN statistics in ONE pass over the
// packed feature rows, instead of bs*(K+2) dense torch.mm calls and six intermediate tensors.
//
// Shape of the work: HBM/L2-bound gathers followed by a tiny per-row
// mat-vec (Cin x Fout, e.g. 20 x 4 at the script default h=2).  Design:
//   * persistent CTAs (grid = multiple of 148 SMs) loop over tiles of TR consecutive rows;
//   * phase 1 (gather): work item = (row, VEC-wide feature chunk); a warp owns 32 consecutive rows
//     of one chunk, so self loads are fully coalesced and CSR segments of neighbouring rows are
//     adjacent; float4 gathers when the feature width allows; results are staged in shared memory
//     as the concatenated x1 tile (the reference's torch.cat never reaches HBM);
//   * phase 2 (mat-vec): thread owns output column o and up to four rows; weights live in shared
//     memory transposed ([Cin][Fout]) so a warp reads consecutive banks; fp32 FMA - tensor cores do
//     not pay at Cin*Fout = 80 (north_star: "only when the feature width makes that linear a real
//     dense contraction");
//   * epilogue: bias, ReLU on outputs >= relu_from, per-feature (sum, sum^2) in registers ->
//     per-CTA partials -> the last CTA reduces them in fixed order and finalises the batch-norm
//     statistics in the same launch.
#include "bn_common.cuh"

int hgnn_grid_cap(int width);

#define SIDE_THREADS 256
#define SIDE_MAX_SMEM (200 * 1024)

template <int VEC> struct Vec;
template <> struct Vec<1> {
    float v;
    __device__ __forceinline__ static Vec load(const float* p) { Vec r; r.v = __ldg(p); return r; }
    __device__ __forceinline__ static Vec zero() { Vec r; r.v = 0.f; return r; }
    __device__ __forceinline__ void fma(float a, const Vec& x) { v = fmaf(a, x.v, v); }
    __device__ __forceinline__ void scale(float a) { v *= a; }
    __device__ __forceinline__ void add(const Vec& o) { v += o.v; }
    __device__ __forceinline__ void warp_reduce() { v = warp_sum(v); }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ void store_scalar(float* p) const { p[0] = v; }
};
template <> struct Vec<4> {
    float4 v;
    __device__ __forceinline__ static Vec load(const float* p) {
        Vec r; r.v = __ldg(reinterpret_cast<const float4*>(p)); return r;
    }
    __device__ __forceinline__ static Vec zero() { Vec r; r.v = make_float4(0.f, 0.f, 0.f, 0.f); return r; }
    __device__ __forceinline__ void fma(float a, const Vec& x) {
        v.x = fmaf(a, x.v.x, v.x); v.y = fmaf(a, x.v.y, v.y);
        v.z = fmaf(a, x.v.z, v.z); v.w = fmaf(a, x.v.w, v.w);
    }
    __device__ __forceinline__ void scale(float a) { v.x *= a; v.y *= a; v.z *= a; v.w *= a; }
    __device__ __forceinline__ void add(const Vec& o) { v.x += o.v.x; v.y += o.v.y; v.z += o.v.z; v.w += o.v.w; }
    __device__ __forceinline__ void warp_reduce() {
        v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
    }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void store_scalar(float* p) const { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
};

static inline double* ws_accum_host(void* ws) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + HGNN_WS_HEADER);
}

// Gather op t of `ops` for row `row`, feature chunk at column offset `xo` of a (.., ldx) matrix.
template <int VEC>
__device__ __forceinline__ Vec<VEC> gather_op(const OpList& ops, int t, int row,
                                              const float* __restrict__ X, int ldx, int xo) {
    const int kind = ops.kind[t];
    if (kind == HGNN_OP_IDENT) return Vec<VEC>::load(X + (size_t)row * ldx + xo);
    if (kind == HGNN_OP_DIAG) {
        Vec<VEC> x = Vec<VEC>::load(X + (size_t)row * ldx + xo);
        x.scale(__ldg(ops.diag[t] + row));
        return x;
    }
    const int* __restrict__ col = ops.col[t];
    const float* __restrict__ val = ops.val[t];
    const int k0 = __ldg(ops.rowptr[t] + row), k1 = __ldg(ops.rowptr[t] + row + 1);
    Vec<VEC> acc = Vec<VEC>::zero();
    // batches of 4: all index/value loads of a batch are issued before the first gather, and all
    // gathers before the first FMA, so a short row costs 3 dependent memory hops (rowptr -> col ->
    // feature row) instead of 2 per entry
    for (int k = k0; k < k1; k += 4) {
        int c[4];
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool on = k + j < k1;
            c[j] = __ldg(col + (on ? k + j : k));     // padded lanes re-read a valid entry, weight 0
            v[j] = on ? __ldg(val + k + j) : 0.f;
        }
        Vec<VEC> x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = Vec<VEC>::load(X + (size_t)c[j] * ldx + xo);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc.fma(v[j], x[j]);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// long rows: the reference's line-graph enumeration bug (functions/operators.py:59) leaves ~E
// phantom line-graph nodes that all link to the edges leaving node 0, so a handful of rows of the
// TRANSPOSED operator hold thousands of entries (2 650 at N=1000) while the median row holds 3.
// A thread that meets a row longer than LONG_ROW defers it; deferred rows are then gathered by a
// whole warp (<= CTA_ROW entries) or the whole CTA, with a fixed-order tree reduction.
// ---------------------------------------------------------------------------------------------
#define LONG_ROW 32
#define CTA_ROW 1024
#define MAX_DEFER 192

struct DeferList {
    int cnt;
    int items[MAX_DEFER];
};

__device__ __forceinline__ int defer_code(int t, int q, int r) { return (t << 24) | (q << 12) | r; }

// Gather op t for (row, chunk) into dst, or push it on the CTA's deferred list when the row is long.
template <int VEC>
__device__ __forceinline__ void gather_or_defer(const OpList& ops, int t, int row,
                                                const float* __restrict__ X, int ldx, int xo,
                                                float* dst, DeferList* dl, int code) {
    if (ops.kind[t] == HGNN_OP_CSR) {
        const int len = __ldg(ops.rowptr[t] + row + 1) - __ldg(ops.rowptr[t] + row);
        if (len > LONG_ROW) {
            const int slot = atomicAdd(&dl->cnt, 1);
            if (slot < MAX_DEFER) {
                dl->items[slot] = code;
                return;
            }
        }
    }
    gather_op<VEC>(ops, t, row, X, ldx, xo).store(dst);
}

template <int VEC>
__device__ __forceinline__ Vec<VEC> strided_gather(const OpList& ops, int t, int row,
                                                   const float* __restrict__ X, int ldx, int xo,
                                                   int first, int stride) {
    const int* __restrict__ col = ops.col[t];
    const float* __restrict__ val = ops.val[t];
    const int k0 = __ldg(ops.rowptr[t] + row), k1 = __ldg(ops.rowptr[t] + row + 1);
    Vec<VEC> a0 = Vec<VEC>::zero(), a1 = Vec<VEC>::zero();
    int k = k0 + first;
    for (; k + stride < k1; k += 2 * stride) {
        Vec<VEC> x0 = Vec<VEC>::load(X + (size_t)__ldg(col + k) * ldx + xo);
        Vec<VEC> x1 = Vec<VEC>::load(X + (size_t)__ldg(col + k + stride) * ldx + xo);
        a0.fma(__ldg(val + k), x0);
        a1.fma(__ldg(val + k + stride), x1);
    }
    if (k < k1) a0.fma(__ldg(val + k), Vec<VEC>::load(X + (size_t)__ldg(col + k) * ldx + xo));
    a0.add(a1);
    return a0;
}

// Process the deferred (row, chunk, op) items of the current tile.  Must be called by ALL threads
// of the CTA after a __syncthreads() that follows the per-thread gather loop.
template <int VEC>
__device__ __forceinline__ void gather_deferred(const OpList& ops, DeferList* dl, int row0,
                                                const float* __restrict__ X, int ldx, int Fblk,
                                                float* tile, int Tp, float* wpart /* [8][4] */) {
    const int nd = min(dl->cnt, MAX_DEFER);
    if (nd == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int it = warp; it < nd; it += nwarps) {          // medium rows: one warp each
        const int code = dl->items[it];
        const int t = code >> 24, q = (code >> 12) & 0xfff, r = code & 0xfff;
        const int row = row0 + r, xo = q * VEC;
        const int len = __ldg(ops.rowptr[t] + row + 1) - __ldg(ops.rowptr[t] + row);
        if (len > CTA_ROW) continue;
        Vec<VEC> acc = strided_gather<VEC>(ops, t, row, X, ldx, xo, lane, 32);
        acc.warp_reduce();
        if (lane == 0) acc.store(tile + r * Tp + t * Fblk + xo);
    }
    for (int it = 0; it < nd; ++it) {                     // very long rows: the whole CTA
        const int code = dl->items[it];
        const int t = code >> 24, q = (code >> 12) & 0xfff, r = code & 0xfff;
        const int row = row0 + r, xo = q * VEC;
        const int len = __ldg(ops.rowptr[t] + row + 1) - __ldg(ops.rowptr[t] + row);
        if (len <= CTA_ROW) continue;                     // uniform across the CTA
        Vec<VEC> acc = strided_gather<VEC>(ops, t, row, X, ldx, xo, threadIdx.x, blockDim.x);
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
}

// Cross-CTA sums: every CTA adds its partial into fp64 accumulators in the workspace
// ([0,256) = ticket, then `width` doubles, zero on entry); the last CTA reads the totals, runs the
// finalizer and zeroes them again for the next launch.
__device__ __forceinline__ double* ws_accum(void* ws) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + HGNN_WS_HEADER);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct SideFwdArgs {
    int R;
    OpList ops;
    const float* Xs; int Fs;
    const int* p_rowptr; const int* p_col; const float* p_pm; const float* p_pd;
    const float* Xc; int Fc;
    const float* Wa; const float* ba; int Ha;
    const float* Wb; const float* bb; int Hb;
    int relu_from;
    float* Z;
    const float* bn_w; const float* bn_b; float* run_mean; float* run_std; float momentum;
    float* stats;
    unsigned int* counter; double* accum;
    int TR, Cin, Cin_pad, Fout;
};

template <int VEC, int VOUT>
__global__ void __launch_bounds__(SIDE_THREADS, 4)
side_fwd_kernel(const SideFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[2 * SIDE_THREADS];   // >= (SIDE_THREADS/32)*32 doubles for the shuffle tree
    __shared__ DeferList dl;
    __shared__ float wpart[32];
    const int Cin = a.Cin, Cp = a.Cin_pad, Fout = a.Fout, TR = a.TR;
    float* Wt = smem;                         // [Cin][Fout]
    float* bias = Wt + Cin * Fout;            // [Fout]
    float* tile = bias + ((Fout + 3) & ~3);   // [TR][Cp]
    const int tid = threadIdx.x;

    for (int i = tid; i < Cin * Fout; i += SIDE_THREADS) {
        const int o = i / Cin, c = i - o * Cin;   // coalesced read of the row-major conv weights
        const float w = (o < a.Ha) ? a.Wa[(size_t)o * Cin + c] : a.Wb[(size_t)(o - a.Ha) * Cin + c];
        Wt[c * Fout + o] = w;
    }
    for (int o = tid; o < Fout; o += SIDE_THREADS)
        bias[o] = (o < a.Ha) ? (a.ba ? a.ba[o] : 0.f) : (a.bb ? a.bb[o - a.Ha] : 0.f);

    const int K = a.ops.n, Fs = a.Fs, Fc = a.Fc;
    const int Qs = Fs / VEC, Qc = Fc / VEC;
    const bool cross = a.p_rowptr != nullptr;
    const int Q = Qs + (cross ? Qc : 0);
    const int xc0 = K * Fs;

    // phase-2 ownership: VOUT outputs of one row per step
    const int NQ = Fout / VOUT;                       // output groups per row
    const int rows_per_pass = SIDE_THREADS / NQ;
    const bool owner = tid < rows_per_pass * NQ;
    const int oq = tid % NQ, rg = tid / NQ;
    float s1[VOUT], s2[VOUT];
#pragma unroll
    for (int j = 0; j < VOUT; ++j) s1[j] = s2[j] = 0.f;
    const int ntiles = (a.R + TR - 1) / TR;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, a.R - row0);
        if (tid == 0) dl.cnt = 0;
        __syncthreads();   // weights ready (first trip) / previous tile consumed
        // ---- phase 1: gather the concatenated x1 rows into shared memory
        for (int i = tid; i < Q * TR; i += SIDE_THREADS) {
            const int q = i / TR, r = i - q * TR;
            if (r >= trc) continue;
            const int row = row0 + r;
            float* trow = tile + r * Cp;
            if (q < Qs) {
                const int xo = q * VEC;
                for (int t = 0; t < K; ++t)
                    gather_or_defer<VEC>(a.ops, t, row, a.Xs, Fs, xo, trow + t * Fs + xo, &dl,
                                         defer_code(t, q, r));
            } else {
                const int xo = (q - Qs) * VEC;
                Vec<VEC> am = Vec<VEC>::zero(), ad = Vec<VEC>::zero();
                const int k0 = __ldg(a.p_rowptr + row), k1 = __ldg(a.p_rowptr + row + 1);
                for (int k = k0; k < k1; k += 4) {      // batched like gather_op
                    int c[4];
                    float vm[4], vd[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool on = k + j < k1;
                        c[j] = __ldg(a.p_col + (on ? k + j : k));
                        vm[j] = on ? __ldg(a.p_pm + k + j) : 0.f;
                        vd[j] = on ? __ldg(a.p_pd + k + j) : 0.f;
                    }
                    Vec<VEC> x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = Vec<VEC>::load(a.Xc + (size_t)c[j] * Fc + xo);
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
        gather_deferred<VEC>(a.ops, &dl, row0, a.Xs, Fs, Fs, tile, Cp, wpart);
        __syncthreads();
        // ---- phase 2: Z = W x1 + b, ReLU, statistics
        if (owner) {
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
    if (a.stats) {
        // per-thread fp32 partial sums (a few rows each) -> fp64 shuffle tree -> binned atomics
        const int nb = hgnn_ws_bins(2 * Fout);
        const bool tree = (32 % NQ) == 0 && (SIDE_THREADS % NQ) == 0;
#pragma unroll
        for (int j = 0; j < VOUT; ++j) {
            double x = 0.0, y = 0.0;
            if (tree) {
                x = cta_reduce_mod(owner ? (double)s1[j] : 0.0, NQ, red);
                y = cta_reduce_mod(owner ? (double)s2[j] : 0.0, NQ, red);
            } else {
                __syncthreads();
                red[tid] = owner ? (double)s1[j] : 0.0;
                red[SIDE_THREADS + tid] = owner ? (double)s2[j] : 0.0;
                __syncthreads();
                if (tid < NQ)
                    for (int k = 0; k < rows_per_pass; ++k) {
                        x += red[k * NQ + tid];
                        y += red[SIDE_THREADS + k * NQ + tid];
                    }
            }
            if (tid < NQ) {
                accum_add(a.accum, 2 * Fout, nb, tid * VOUT + j, x);
                accum_add(a.accum, 2 * Fout, nb, Fout + tid * VOUT + j, y);
            }
        }
        if (last_block_ticket(a.counter)) {
            bn_finalize_accum(a.accum, Fout, a.R, a.bn_w, a.bn_b, a.run_mean, a.run_std, a.momentum, a.stats);
            if (tid == 0) *a.counter = 0;
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pad_stride(int width, int vec) {
    if (vec == 4) {
        int p = (width + 3) & ~3;
        if (((p >> 2) & 1) == 0) p += 4;   // stride/4 odd: 8 consecutive rows hit 8 distinct bank groups
        return p;
    }
    return width | 1;
}

// CTAs of `kernel` resident on the whole GPU with `smem` bytes of dynamic shared memory.  Cached per
// (kernel address, smem); also raises the kernel's opt-in shared-memory limit when needed.  (All
// instantiations of a kernel template share one function-pointer TYPE, so the cache must be keyed by
// the pointer VALUE.)
struct SideOccEntry { const void* fn; size_t smem; int occ; };
static int side_resident_impl(const void* fn, size_t smem, int threads) {
    static SideOccEntry cache[64];
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
static int resident_ctas(K kernel, size_t smem) {
    return side_resident_impl(reinterpret_cast<const void*>(kernel), smem, SIDE_THREADS);
}

extern "C" int hgnn_side_fwd(const hgnn_side_t* side, const float* Wa, const float* ba, int Ha,
                             const float* Wb, const float* bb, int Hb, int relu_from, float* Z,
                             const float* bn_weight, const float* bn_bias, float* running_mean,
                             float* running_std, float momentum, float* stats, void* ws,
                             long long ws_bytes, hgnn_stream_t stream) {
    HGNN_REQUIRE(side && Z, "null argument");
    SideFwdArgs a;
    HGNN_REQUIRE(make_oplist(side->ops, side->n_ops, &a.ops) == 0 && side->n_ops >= 1,
                 "bad operator list");
    a.R = side->R;
    a.Xs = side->Xs; a.Fs = side->Fs;
    a.p_rowptr = side->p_rowptr; a.p_col = side->p_col; a.p_pm = side->p_pm; a.p_pd = side->p_pd;
    a.Xc = side->Xc; a.Fc = side->p_rowptr ? side->Fc : 0;
    a.Wa = Wa; a.ba = ba; a.Ha = Ha; a.Wb = Wb; a.bb = bb; a.Hb = Hb;
    a.relu_from = relu_from; a.Z = Z;
    a.bn_w = bn_weight; a.bn_b = bn_bias; a.run_mean = running_mean; a.run_std = running_std;
    a.momentum = momentum; a.stats = stats;
    HGNN_REQUIRE(a.R >= 0 && a.Fs >= 1 && a.Xs, "bad self features");
    HGNN_REQUIRE(!side->p_rowptr || (side->Fc >= 1 && side->Xc && side->p_col && side->p_pm && side->p_pd),
                 "bad cross part");
    HGNN_REQUIRE(Ha >= 0 && Hb >= 0 && Ha + Hb >= 1 && Ha + Hb <= SIDE_THREADS, "bad output width");
    HGNN_REQUIRE(!stats || Ha + Hb <= 128, "batch-norm statistics support at most 128 features");
    HGNN_REQUIRE((Ha == 0 || Wa) && (Hb == 0 || Wb), "null weights");
    if (a.R == 0) return HGNN_OK;
    a.Fout = Ha + Hb;
    a.Cin = side->n_ops * a.Fs + 2 * a.Fc;
    const bool vec4 = (a.Fs % 4 == 0) && (a.Fc % 4 == 0) && aligned16(a.Xs) && (a.Fc == 0 || aligned16(a.Xc));
    const bool vout4 = vec4 && (a.Fout % 4 == 0) && aligned16(Z);
    a.Cin_pad = pad_stride(a.Cin, vec4 ? 4 : 1);
    const int rows_per_pass = SIDE_THREADS / (vout4 ? a.Fout / 4 : a.Fout);
    const int items_per_row = (a.Fs + a.Fc) / (vec4 ? 4 : 1);
    int TR = min(rows_per_pass, max(32, SIDE_THREADS / max(1, items_per_row)));   // ~1 gather item per thread
    size_t fixed = ((size_t)a.Cin * a.Fout + ((a.Fout + 3) & ~3)) * sizeof(float);
    while (TR > 1 && fixed + (size_t)TR * a.Cin_pad * sizeof(float) > SIDE_MAX_SMEM) TR >>= 1;
    size_t smem = fixed + (size_t)TR * a.Cin_pad * sizeof(float);
    if (smem > SIDE_MAX_SMEM) {
        hgnn_set_error("hgnn_side_fwd: Cin=%d x Fout=%d does not fit shared memory", a.Cin, a.Fout);
        return HGNN_ERR_ARG;
    }
    a.TR = TR;
    a.counter = nullptr; a.accum = nullptr;
    if (stats) {
        HGNN_REQUIRE(ws, "workspace required for batch-norm statistics");
        if (ws_bytes < hgnn_workspace_bytes(2 * a.Fout)) {
            hgnn_set_error("hgnn_side_fwd: workspace too small");
            return HGNN_ERR_WORKSPACE;
        }
        a.counter = (unsigned int*)ws;
        a.accum = ws_accum_host(ws);
    }
    const int ntiles = ceil_div(a.R, TR);
    cudaStream_t s = to_stream(stream);
    if (vec4 && vout4) {
        int grid = balanced_grid(ntiles, resident_ctas(side_fwd_kernel<4, 4>, smem));
        side_fwd_kernel<4, 4><<<grid, SIDE_THREADS, smem, s>>>(a);
    } else if (vec4) {
        int grid = balanced_grid(ntiles, resident_ctas(side_fwd_kernel<4, 1>, smem));
        side_fwd_kernel<4, 1><<<grid, SIDE_THREADS, smem, s>>>(a);
    } else {
        int grid = balanced_grid(ntiles, resident_ctas(side_fwd_kernel<1, 1>, smem));
        side_fwd_kernel<1, 1><<<grid, SIDE_THREADS, smem, s>>>(a);
    }
    return hgnn_check_launch("hgnn_side_fwd");
}

// ---------------------------------------------------------------------------------------------
// backward gather:  T = [opsT_t G]_t ;  gX (+)= W_blocks^T T ;  dW_blocks = sum_rows T (x) X
// ---------------------------------------------------------------------------------------------
struct SideBwdArgs {
    int R;
    OpList ops;
    const float* G; int Fg;
    const float* X; int Fx;
    const float* Wa; int Ha; const float* Wb; int Hb;
    int Cin, col0;
    float* gX; int accumulate;
    float* dWa; float* dWb;
    unsigned int* counter; double* accum;
    int TR, nT, Tp, Xp, NG, P;
};

template <int VEC, int VOUT>
__global__ void __launch_bounds__(SIDE_THREADS, 4)
side_bwd_kernel(const SideBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ DeferList dl;
    __shared__ float wpart[32];
    const int nT = a.nT, Tp = a.Tp, Fx = a.Fx, Xp = a.Xp, Fg = a.Fg, TR = a.TR, P = a.P, NG = a.NG;
    float* Wsm = smem;                                  // [nT][Fx]  (W blocks, transposed view)
    float* tile = Wsm + ((nT * Fx + 3) & ~3);           // [TR][Tp]
    float* xt = tile + TR * Tp;                         // [TR][Xp]
    float* dacc = xt + ((TR * Xp + 3) & ~3);            // [NG][P]
    const int tid = threadIdx.x;
    const int K = a.ops.n;
    const bool want_dw = a.dWa || a.dWb;

    for (int i = tid; i < nT * Fx; i += SIDE_THREADS) {
        const int c = i / Fx, f = i - c * Fx;
        const int t = c / Fg, o = c - t * Fg;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Wsm[i] = wrow[a.col0 + t * Fx + f];
    }
    for (int i = tid; i < NG * P; i += SIDE_THREADS) dacc[i] = 0.f;

    const int Q = Fg / VEC;
    const int NQ = Fx / VOUT;                           // gX groups per row
    const int rows_per_pass = SIDE_THREADS / NQ;
    const bool owner = tid < rows_per_pass * NQ;
    const int fq = tid % NQ, rg = tid / NQ;
    // dW ownership: slot = (group g, column c, feature group) ; VOUT features per slot
    const int PS = nT * NQ;                             // slots per row group
    const int ntiles = (a.R + TR - 1) / TR;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, a.R - row0);
        if (tid == 0) dl.cnt = 0;
        __syncthreads();
        // ---- phase 1: transposed gather of G into the T tile; stage the rows' own features
        for (int i = tid; i < Q * TR; i += SIDE_THREADS) {
            const int q = i / TR, r = i - q * TR;
            if (r >= trc) continue;
            const int xo = q * VEC;
            float* trow = tile + r * Tp;
            for (int t = 0; t < K; ++t)
                gather_or_defer<VEC>(a.ops, t, row0 + r, a.G, Fg, xo, trow + t * Fg + xo, &dl,
                                     defer_code(t, q, r));
        }
        if (VOUT == 4) {
            for (int i = tid; i < trc * NQ; i += SIDE_THREADS) {
                const int r = i / NQ, g4 = i - r * NQ;
                *reinterpret_cast<float4*>(xt + r * Xp + g4 * 4) =
                    __ldg(reinterpret_cast<const float4*>(a.X + (size_t)row0 * Fx) + i);
            }
        } else {
            for (int i = tid; i < trc * Fx; i += SIDE_THREADS) {
                const int r = i / Fx, f = i - r * Fx;
                xt[r * Xp + f] = a.X[(size_t)row0 * Fx + i];
            }
        }
        __syncthreads();
        gather_deferred<VEC>(a.ops, &dl, row0, a.G, Fg, Fg, tile, Tp, wpart);
        __syncthreads();
        // ---- phase 2: gX = W^T T   (VOUT features of one row per step)
        if (a.gX && owner) {
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
                float* dst = a.gX + (size_t)(row0 + r) * Fx + fq * VOUT;
                if (VOUT == 4) {
                    float4 o = make_float4(acc[0], acc[1 % VOUT], acc[2 % VOUT], acc[3 % VOUT]);
                    if (a.accumulate) {
                        const float4 old = *reinterpret_cast<const float4*>(dst);
                        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                    }
                    *reinterpret_cast<float4*>(dst) = o;
                } else {
                    dst[0] = a.accumulate ? dst[0] + acc[0] : acc[0];
                }
            }
        }
        // ---- phase 3: dW[c][f] += sum_r T[r][c] * X[r][f]; every (group, c, feature group) slot has
        //      one owner thread, so the shared accumulators need no atomics
        if (want_dw) {
            for (int sidx = tid; sidx < NG * PS; sidx += SIDE_THREADS) {
                const int g = sidx / PS, p = sidx - g * PS;
                const int c = p / NQ, g4 = p - c * NQ;
                float acc[VOUT];
#pragma unroll
                for (int j = 0; j < VOUT; ++j) acc[j] = 0.f;
                for (int r = g; r < trc; r += NG) {
                    const float tv = tile[r * Tp + c];
                    if (VOUT == 4) {
                        const float4 x = *reinterpret_cast<const float4*>(xt + r * Xp + g4 * 4);
                        acc[0] = fmaf(tv, x.x, acc[0]);
                        acc[1 % VOUT] = fmaf(tv, x.y, acc[1 % VOUT]);
                        acc[2 % VOUT] = fmaf(tv, x.z, acc[2 % VOUT]);
                        acc[3 % VOUT] = fmaf(tv, x.w, acc[3 % VOUT]);
                    } else {
                        acc[0] = fmaf(tv, xt[r * Xp + g4], acc[0]);
                    }
                }
                float* d = dacc + (size_t)g * P + c * Fx + g4 * VOUT;
#pragma unroll
                for (int j = 0; j < VOUT; ++j) d[j] += acc[j];
            }
        }
    }
    if (want_dw) {
        __syncthreads();
        for (int p = tid; p < P; p += SIDE_THREADS) {
            float acc = 0.f;
            for (int g = 0; g < NG; ++g) acc += dacc[g * P + p];
            accum_add(a.accum, P, hgnn_ws_bins(P), p, (double)acc);
        }
        if (last_block_ticket(a.counter)) {
            const int nb = hgnn_ws_bins(P);
            for (int p = tid; p < P; p += SIDE_THREADS) {
                const float acc = (float)accum_take(a.accum, P, nb, p);
                const int c = p / Fx, f = p - c * Fx;
                const int t = c / Fg, o = c - t * Fg;
                float* drow = (o < a.Ha) ? a.dWa + (size_t)o * a.Cin : a.dWb + (size_t)(o - a.Ha) * a.Cin;
                drow[a.col0 + t * Fx + f] = acc;
            }
            if (tid == 0) *a.counter = 0;
        }
    }
}

extern "C" int hgnn_side_bwd_gather(const hgnn_op_t* opsT, int n_ops, int R, const float* G, int Fg,
                                    const float* X, int Fx, const float* Wa, int Ha, const float* Wb,
                                    int Hb, int Cin, int col0, float* gX, int accumulate, float* dWa,
                                    float* dWb, void* ws, long long ws_bytes, hgnn_stream_t stream) {
    SideBwdArgs a;
    HGNN_REQUIRE(opsT && n_ops >= 1 && make_oplist(opsT, n_ops, &a.ops) == 0, "bad operator list");
    HGNN_REQUIRE(G && X && Fg >= 1 && Fx >= 1 && Fx <= SIDE_THREADS && R >= 0, "bad argument");
    HGNN_REQUIRE(Ha >= 0 && Hb >= 0 && Ha + Hb == Fg, "Ha + Hb must equal the width of G");
    HGNN_REQUIRE((Ha == 0 || Wa) && (Hb == 0 || Wb), "null weights");
    HGNN_REQUIRE(col0 >= 0 && col0 + n_ops * Fx <= Cin, "column block out of range");
    const bool want_dw = dWa || dWb;
    HGNN_REQUIRE(!want_dw || ((Ha == 0 || dWa) && (Hb == 0 || dWb)), "dWa/dWb must both be given");
    if (R == 0) return HGNN_OK;   // caller zero-fills dW when there are no rows
    a.R = R; a.G = G; a.Fg = Fg; a.X = X; a.Fx = Fx;
    a.Wa = Wa; a.Ha = Ha; a.Wb = Wb; a.Hb = Hb; a.Cin = Cin; a.col0 = col0;
    a.gX = gX; a.accumulate = accumulate; a.dWa = dWa; a.dWb = dWb;
    a.nT = n_ops * Fg;
    a.P = a.nT * Fx;
    const bool vec4 = (Fg % 4 == 0) && aligned16(G);
    const bool vout4 = vec4 && (Fx % 4 == 0) && aligned16(X) && (!gX || aligned16(gX));
    const int slots = a.nT * (vout4 ? Fx / 4 : Fx);
    a.NG = slots >= SIDE_THREADS ? 1 : SIDE_THREADS / slots;
    a.Tp = pad_stride(a.nT, vec4 ? 4 : 1);
    a.Xp = vout4 ? pad_stride(Fx, 4) : (Fx | 1);
    const int rows_per_pass = SIDE_THREADS / (vout4 ? Fx / 4 : Fx);
    int TR = min(rows_per_pass, max(32, SIDE_THREADS / max(1, Fg / (vec4 ? 4 : 1))));
    auto smem_for = [&](int tr) {
        return ((size_t)((a.nT * Fx + 3) & ~3) + (size_t)tr * a.Tp + (size_t)((tr * a.Xp + 3) & ~3) +
                (size_t)a.NG * a.P) * sizeof(float);
    };
    while (TR > 1 && smem_for(TR) > SIDE_MAX_SMEM) TR >>= 1;
    size_t smem = smem_for(TR);
    if (smem > SIDE_MAX_SMEM) {
        hgnn_set_error("hgnn_side_bwd_gather: %d x %d weight block does not fit shared memory", a.nT, Fx);
        return HGNN_ERR_ARG;
    }
    a.TR = TR;
    a.counter = nullptr; a.accum = nullptr;
    if (want_dw) {
        HGNN_REQUIRE(ws, "workspace required for dW");
        if (ws_bytes < hgnn_workspace_bytes(a.P)) {
            hgnn_set_error("hgnn_side_bwd_gather: workspace too small");
            return HGNN_ERR_WORKSPACE;
        }
        a.counter = (unsigned int*)ws;
        a.accum = ws_accum_host(ws);
    }
    const int ntiles = ceil_div(R, TR);
    cudaStream_t s = to_stream(stream);
    if (vec4 && vout4) {
        int grid = balanced_grid(ntiles, resident_ctas(side_bwd_kernel<4, 4>, smem));
        side_bwd_kernel<4, 4><<<grid, SIDE_THREADS, smem, s>>>(a);
    } else if (vec4) {
        int grid = balanced_grid(ntiles, resident_ctas(side_bwd_kernel<4, 1>, smem));
        side_bwd_kernel<4, 1><<<grid, SIDE_THREADS, smem, s>>>(a);
    } else {
        int grid = balanced_grid(ntiles, resident_ctas(side_bwd_kernel<1, 1>, smem));
        side_bwd_kernel<1, 1><<<grid, SIDE_THREADS, smem, s>>>(a);
    }
    return hgnn_check_launch("hgnn_side_bwd_gather");
}
