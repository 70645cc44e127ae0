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
    __device__ __forceinline__ void store(float* p) const { *p = v; }
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
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
};

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
    int k = k0;
    for (; k + 1 < k1; k += 2) {  // two independent gathers in flight per lane
        const int c0 = __ldg(col + k), c1 = __ldg(col + k + 1);
        const float v0 = __ldg(val + k), v1 = __ldg(val + k + 1);
        Vec<VEC> x0 = Vec<VEC>::load(X + (size_t)c0 * ldx + xo);
        Vec<VEC> x1 = Vec<VEC>::load(X + (size_t)c1 * ldx + xo);
        acc.fma(v0, x0);
        acc.fma(v1, x1);
    }
    if (k < k1) {
        Vec<VEC> x0 = Vec<VEC>::load(X + (size_t)__ldg(col + k) * ldx + xo);
        acc.fma(__ldg(val + k), x0);
    }
    return acc;
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
    unsigned int* counter; double* partial;
    int TR, Cin, Cin_pad, Fout;
};

template <int VEC>
__global__ void __launch_bounds__(SIDE_THREADS)
side_fwd_kernel(const SideFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double red[2 * SIDE_THREADS];
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

    const int rows_per_pass = SIDE_THREADS / Fout;
    const bool owner = tid < rows_per_pass * Fout;
    const int o = tid % Fout, rg = tid / Fout;
    double s1 = 0.0, s2 = 0.0;
    const int ntiles = (a.R + TR - 1) / TR;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, a.R - row0);
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
                    gather_op<VEC>(a.ops, t, row, a.Xs, Fs, xo).store(trow + t * Fs + xo);
            } else {
                const int xo = (q - Qs) * VEC;
                Vec<VEC> am = Vec<VEC>::zero(), ad = Vec<VEC>::zero();
                const int k0 = __ldg(a.p_rowptr + row), k1 = __ldg(a.p_rowptr + row + 1);
                for (int k = k0; k < k1; ++k) {
                    Vec<VEC> x = Vec<VEC>::load(a.Xc + (size_t)__ldg(a.p_col + k) * Fc + xo);
                    am.fma(__ldg(a.p_pm + k), x);
                    ad.fma(__ldg(a.p_pd + k), x);
                }
                am.store(trow + xc0 + xo);
                ad.store(trow + xc0 + Fc + xo);
            }
        }
        __syncthreads();
        // ---- phase 2: Z = W x1 + b, ReLU, statistics
        if (owner) {
            for (int rb = rg; rb < trc; rb += 4 * rows_per_pass) {
                float acc0 = bias[o], acc1 = acc0, acc2 = acc0, acc3 = acc0;
                const float* t0 = tile + min(rb, TR - 1) * Cp;
                const float* t1 = tile + min(rb + rows_per_pass, TR - 1) * Cp;
                const float* t2 = tile + min(rb + 2 * rows_per_pass, TR - 1) * Cp;
                const float* t3 = tile + min(rb + 3 * rows_per_pass, TR - 1) * Cp;
                const float* w = Wt + o;
#pragma unroll 4
                for (int c = 0; c < Cin; ++c) {
                    const float wv = w[c * Fout];
                    acc0 = fmaf(t0[c], wv, acc0);
                    acc1 = fmaf(t1[c], wv, acc1);
                    acc2 = fmaf(t2[c], wv, acc2);
                    acc3 = fmaf(t3[c], wv, acc3);
                }
                float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = rb + j * rows_per_pass;
                    if (r < trc) {
                        float v = accs[j];
                        if (o >= a.relu_from) v = fmaxf(v, 0.f);
                        a.Z[(size_t)(row0 + r) * Fout + o] = v;
                        s1 += (double)v;
                        s2 += (double)v * (double)v;
                    }
                }
            }
        }
    }
    if (a.stats) {
        __syncthreads();
        ColOwner co(Fout, SIDE_THREADS);
        cta_column_partials(s1, s2, Fout, co, red, a.partial);
        if (last_block_ticket(a.counter)) {
            bn_finalize(a.partial, gridDim.x, Fout, a.R, a.bn_w, a.bn_b, a.run_mean, a.run_std,
                        a.momentum, a.stats);
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
    HGNN_REQUIRE((Ha == 0 || Wa) && (Hb == 0 || Wb), "null weights");
    if (a.R == 0) return HGNN_OK;
    a.Fout = Ha + Hb;
    a.Cin = side->n_ops * a.Fs + 2 * a.Fc;
    const bool vec4 = (a.Fs % 4 == 0) && (a.Fc % 4 == 0) && aligned16(a.Xs) && (a.Fc == 0 || aligned16(a.Xc));
    a.Cin_pad = pad_stride(a.Cin, vec4 ? 4 : 1);
    const int rows_per_pass = SIDE_THREADS / a.Fout;
    int TR = min(256, 4 * rows_per_pass);
    size_t fixed = ((size_t)a.Cin * a.Fout + ((a.Fout + 3) & ~3)) * sizeof(float);
    while (TR > 1 && fixed + (size_t)TR * a.Cin_pad * sizeof(float) > SIDE_MAX_SMEM) TR >>= 1;
    size_t smem = fixed + (size_t)TR * a.Cin_pad * sizeof(float);
    if (smem > SIDE_MAX_SMEM) {
        hgnn_set_error("hgnn_side_fwd: Cin=%d x Fout=%d does not fit shared memory", a.Cin, a.Fout);
        return HGNN_ERR_ARG;
    }
    a.TR = TR;
    a.counter = nullptr; a.partial = nullptr;
    int cap = HGNN_MAX_GRID;
    if (stats) {
        HGNN_REQUIRE(ws, "workspace required for batch-norm statistics");
        if (ws_bytes < hgnn_workspace_bytes(2 * a.Fout)) {
            hgnn_set_error("hgnn_side_fwd: workspace too small");
            return HGNN_ERR_WORKSPACE;
        }
        a.counter = (unsigned int*)ws;
        a.partial = (double*)((char*)ws + HGNN_WS_HEADER);
        cap = hgnn_grid_cap(2 * a.Fout);
    }
    const int ntiles = ceil_div(a.R, TR);
    int ctas_per_sm = (int)min((size_t)8, (size_t)(220 * 1024) / (smem + 4096 + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int grid = min(min(ntiles, HGNN_SM_COUNT * ctas_per_sm), cap);
    cudaStream_t s = to_stream(stream);
    if (vec4) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(side_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        side_fwd_kernel<4><<<grid, SIDE_THREADS, smem, s>>>(a);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(side_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        side_fwd_kernel<1><<<grid, SIDE_THREADS, smem, s>>>(a);
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
    unsigned int* counter; float* partial;
    int TR, nT, Tp, Xp, NG, P;
};

template <int VEC>
__global__ void __launch_bounds__(SIDE_THREADS)
side_bwd_kernel(const SideBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int nT = a.nT, Tp = a.Tp, Fx = a.Fx, Xp = a.Xp, Fg = a.Fg, TR = a.TR, P = a.P, NG = a.NG;
    float* Wsm = smem;                                  // [nT][Fx]  (W blocks, transposed view)
    float* tile = Wsm + ((nT * Fx + 3) & ~3);           // [TR][Tp]
    float* xt = tile + TR * Tp;                         // [TR][Xp]
    float* dacc = xt + ((TR * Xp + 3) & ~3);            // [NG][P]
    const int tid = threadIdx.x;
    const int K = a.ops.n;

    for (int i = tid; i < nT * Fx; i += SIDE_THREADS) {
        const int c = i / Fx, f = i - c * Fx;
        const int t = c / Fg, o = c - t * Fg;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Wsm[i] = wrow[a.col0 + t * Fx + f];
    }
    for (int i = tid; i < NG * P; i += SIDE_THREADS) dacc[i] = 0.f;

    const int Q = Fg / VEC;
    const int rows_per_pass = SIDE_THREADS / Fx;
    const bool owner = tid < rows_per_pass * Fx;
    const int fo = tid % Fx, rg = tid / Fx;
    const int ntiles = (a.R + TR - 1) / TR;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, a.R - row0);
        __syncthreads();
        // ---- phase 1: transposed gather of G into the T tile; stage the rows' own features
        for (int i = tid; i < Q * TR; i += SIDE_THREADS) {
            const int q = i / TR, r = i - q * TR;
            if (r >= trc) continue;
            const int xo = q * VEC;
            float* trow = tile + r * Tp;
            for (int t = 0; t < K; ++t)
                gather_op<VEC>(a.ops, t, row0 + r, a.G, Fg, xo).store(trow + t * Fg + xo);
        }
        for (int i = tid; i < trc * Fx; i += SIDE_THREADS) {
            const int r = i / Fx, f = i - r * Fx;
            xt[r * Xp + f] = a.X[(size_t)row0 * Fx + i];
        }
        __syncthreads();
        // ---- phase 2: gX = W^T T
        if (a.gX && owner) {
            for (int rb = rg; rb < trc; rb += 4 * rows_per_pass) {
                float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
                const float* t0 = tile + min(rb, TR - 1) * Tp;
                const float* t1 = tile + min(rb + rows_per_pass, TR - 1) * Tp;
                const float* t2 = tile + min(rb + 2 * rows_per_pass, TR - 1) * Tp;
                const float* t3 = tile + min(rb + 3 * rows_per_pass, TR - 1) * Tp;
                const float* w = Wsm + fo;
#pragma unroll 4
                for (int c = 0; c < nT; ++c) {
                    const float wv = w[c * Fx];
                    acc0 = fmaf(t0[c], wv, acc0);
                    acc1 = fmaf(t1[c], wv, acc1);
                    acc2 = fmaf(t2[c], wv, acc2);
                    acc3 = fmaf(t3[c], wv, acc3);
                }
                float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = rb + j * rows_per_pass;
                    if (r < trc) {
                        float* dst = a.gX + (size_t)(row0 + r) * Fx + fo;
                        *dst = a.accumulate ? (*dst + accs[j]) : accs[j];
                    }
                }
            }
        }
        // ---- phase 3: dW[c][f] += sum_r T[r][c] * X[r][f]; every (group, pair) slot has one owner
        if (a.dWa || a.dWb) {
            for (int s = tid; s < NG * P; s += SIDE_THREADS) {
                const int g = s / P, p = s - g * P;
                const int c = p / Fx, f = p - c * Fx;
                float acc = 0.f;
                for (int r = g; r < trc; r += NG) acc = fmaf(tile[r * Tp + c], xt[r * Xp + f], acc);
                dacc[s] += acc;
            }
        }
    }
    if (a.dWa || a.dWb) {
        __syncthreads();
        for (int p = tid; p < P; p += SIDE_THREADS) {
            float acc = 0.f;
            for (int g = 0; g < NG; ++g) acc += dacc[g * P + p];
            a.partial[(size_t)blockIdx.x * P + p] = acc;
        }
        if (last_block_ticket(a.counter)) {
            for (int p = tid; p < P; p += SIDE_THREADS) {
                float acc = 0.f;
                for (int b = 0; b < (int)gridDim.x; ++b) acc += a.partial[(size_t)b * P + p];
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
    a.NG = a.P >= SIDE_THREADS ? 1 : SIDE_THREADS / a.P;
    const bool vec4 = (Fg % 4 == 0) && aligned16(G);
    a.Tp = pad_stride(a.nT, vec4 ? 4 : 1);
    a.Xp = Fx | 1;
    const int rows_per_pass = SIDE_THREADS / Fx;
    int TR = min(256, 4 * rows_per_pass);
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
    a.counter = nullptr; a.partial = nullptr;
    int cap = HGNN_MAX_GRID;
    if (want_dw) {
        HGNN_REQUIRE(ws, "workspace required for dW");
        if (ws_bytes < hgnn_workspace_bytes(a.P)) {
            hgnn_set_error("hgnn_side_bwd_gather: workspace too small");
            return HGNN_ERR_WORKSPACE;
        }
        a.counter = (unsigned int*)ws;
        a.partial = (float*)((char*)ws + HGNN_WS_HEADER);
        cap = hgnn_grid_cap(a.P);
    }
    const int ntiles = ceil_div(R, TR);
    int ctas_per_sm = (int)min((size_t)8, (size_t)(220 * 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int grid = min(min(ntiles, HGNN_SM_COUNT * ctas_per_sm), cap);
    cudaStream_t s = to_stream(stream);
    if (vec4) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(side_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        side_bwd_kernel<4><<<grid, SIDE_THREADS, smem, s>>>(a);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(side_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        side_bwd_kernel<1><<<grid, SIDE_THREADS, smem, s>>>(a);
    }
    return hgnn_check_launch("hgnn_side_bwd_gather");
}
