// engine_rowg.cuh -- thread-per-row engine kernels for the ODD small widths around the width-4 states:
// layer 0 of a model (node inputs of dim_input = 5 features, the single line-graph degree feature) and
// the readout (dim_output = 1 or 2 outputs).  Same structure as engine_row4.cuh - one thread owns one
// row, phase 0 (parameters + graph structure) before griddepcontrol.wait, phase 1 (accumulators +
// feature rows) after it - with the widths as template parameters instead of float4.  Six launches of
// a step ran on the generic tile kernels before and took 15 % of it (profiles/README.md).
// Reference: models/layers/layers_mnb.py:52-69,88-95,189-225,379-388 (layer_simple, layer_last,
// layer_with_lg_1, layer_last_lg) with the feature maps of models/gnns/model_mnb.py:48-50,98-100.
// Included by engine.cu inside namespace eng, after engine_row4.cuh.
#pragma once

template <int F>
struct RowV {
    float v[F];
};

template <int F>
__device__ __forceinline__ RowV<F> ldrow(const float* __restrict__ X, int row) {
    RowV<F> r;
    if constexpr (F == 4) {
        const float4 t = ld4(X + (size_t)row * 4);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
#pragma unroll
        for (int f = 0; f < F; ++f) r.v[f] = __ldg(X + (size_t)row * F + f);
    }
    return r;
}

// scale / shift (and mean / rstd) of a width-F input: identity unless F == 4 and the tensor is normalised
template <int F>
struct BnV {
    float sc[F], sh[F], mu[F], rs[F];
    __device__ __forceinline__ void identity() {
#pragma unroll
        for (int f = 0; f < F; ++f) { sc[f] = 1.f; sh[f] = 0.f; mu[f] = 0.f; rs[f] = 1.f; }
    }
    __device__ __forceinline__ void from4(const Bn4& b) {
        const float s4[4] = {b.sc.x, b.sc.y, b.sc.z, b.sc.w}, h4[4] = {b.sh.x, b.sh.y, b.sh.z, b.sh.w};
        const float m4[4] = {b.mu.x, b.mu.y, b.mu.z, b.mu.w}, r4[4] = {b.rs.x, b.rs.y, b.rs.z, b.rs.w};
#pragma unroll
        for (int f = 0; f < F; ++f) { sc[f] = s4[f & 3]; sh[f] = h4[f & 3]; mu[f] = m4[f & 3]; rs[f] = r4[f & 3]; }
    }
};

template <int B, bool TWO, int F>
struct GatherBatchG {
    int c[B];
    float v[B], v2[TWO ? B : 1];
    RowV<F> x[B];
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
        for (int j = 0; j < B; ++j) {
            if (c[j] >= 0) {
                x[j] = ldrow<F>(X, c[j]);
            } else {
#pragma unroll
                for (int f = 0; f < F; ++f) x[j].v[f] = 0.f;
            }
        }
    }
    __device__ __forceinline__ void accumulate(float (&acc)[F], float& ws, float (&acc2)[F], float& ws2) const {
#pragma unroll
        for (int j = 0; j < B; ++j) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                acc[f] = fmaf(v[j], x[j].v[f], acc[f]);
                if (TWO) acc2[f] = fmaf(v2[j], x[j].v[f], acc2[f]);
            }
            ws += v[j];
            if (TWO) ws2 += v2[j];
        }
    }
};

// Forward of one side with self width FS, cross width FC (ignored unless CROSS) and FO <= 4 outputs.
// Uses Fwd4Args (engine_row4.cuh); Z rows are FO floats wide; the statistics (acc_out) need FO == 4.
template <int NCSR, bool CROSS, int BA, int BP, int FS, int FC, int FO>
__global__ void __launch_bounds__(R4_THREADS)
fwd_rowg_kernel(const Fwd4Args a) {
    constexpr int NS = 2 + NCSR;                              // self blocks
    constexpr int CIN = NS * FS + (CROSS ? 2 * FC : 0);
    constexpr int FCC = CROSS ? FC : 1;
    __shared__ float W[4 * CIN];                               // [o][Cin], rows >= FO are zero
    __shared__ float bias[4];
    __shared__ double red[(R4_THREADS / 32) * 8];
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    // ---- phase 0: parameters and graph structure
    for (int i = tid; i < 4 * CIN; i += R4_THREADS) {
        const int o = i / CIN, c = i - o * CIN;
        float w = 0.f;
        if (o < a.Ha) w = a.Wa[(size_t)o * CIN + c];
        else if (o < a.Ha + a.Hb) w = a.Wb[(size_t)(o - a.Ha) * CIN + c];
        W[i] = w;
    }
    if (tid < 4) {
        float b = 0.f;
        if (tid < a.Ha) b = a.ba ? a.ba[tid] : 0.f;
        else if (tid < a.Ha + a.Hb) b = a.bb ? a.bb[tid - a.Ha] : 0.f;
        bias[tid] = b;
    }
    Bn4Loader ls, lc;
    ls.issue_params(a.bn_s);
    if (CROSS) lc.issue_params(a.bn_c);
    const int stride = gridDim.x * R4_THREADS;
    int row = blockIdx.x * R4_THREADS + tid;
    float d = 0.f, rw = 1.f;
    int rr = 0;                                                // the row behind list entry `row`
    int k0[NCSR], k1[NCSR], p0 = 0, p1 = 0;
    GatherBatchG<BA, false, FS> ga;
    GatherBatchG<BP, true, FCC> gb;
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
        if (rw <= 0.f) { k1[0] = k0[0]; p1 = p0; }
        ga.load_entries(a.col[0], a.val[0], nullptr, k0[0], k1[0]);
        if (CROSS) gb.load_entries(a.p_col, a.p_pm, a.p_pd, p0, p1);
    };
    if (row < a.R) load_structure(row);
    // ---- phase 1: everything the producer wrote
    pdl_wait();
    ls.issue_acc(a.bn_s);
    if (CROSS) lc.issue_acc(a.bn_c);
    RowV<FS> xs_raw;
#pragma unroll
    for (int f = 0; f < FS; ++f) xs_raw.v[f] = 0.f;
    if (row < a.R) {
        xs_raw = ldrow<FS>(a.Xs, rr);
        ga.load_rows(a.Xs);
        if (CROSS) gb.load_rows(a.Xc);
    }
    BnV<FS> bs;
    BnV<FCC> bc;
    bs.identity();
    bc.identity();
    if (FS == 4 && ls.mode != 0) bs.from4(ls.resolve(a.bn_s));
    if (CROSS && FC == 4 && lc.mode != 0) bc.from4(lc.resolve(a.bn_c));
    __syncthreads();                                           // weights in shared memory
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};

    for (bool first = true; row < a.R; row += stride, first = false) {
        if (!first) {
            load_structure(row);
            xs_raw = ldrow<FS>(a.Xs, rr);
            ga.load_rows(a.Xs);
            if (CROSS) gb.load_rows(a.Xc);
        }
        if (rw <= 0.f) continue;                                // a skipped copy of a phantom line-graph row
        float x1[CIN];
#pragma unroll
        for (int f = 0; f < FS; ++f) {
            const float xn = fmaf(xs_raw.v[f], bs.sc[f], bs.sh[f]);
            x1[f] = xn;
            x1[FS + f] = d * xn;
        }
        // sum val*(s*z+t) = s*(sum val*z) + t*(sum val)
        float acc0[FS], am[FCC], ad[FCC], unused_s[FS];
#pragma unroll
        for (int f = 0; f < FS; ++f) { acc0[f] = 0.f; unused_s[f] = 0.f; }
#pragma unroll
        for (int f = 0; f < FCC; ++f) { am[f] = 0.f; ad[f] = 0.f; }
        float ws0 = 0.f, wm = 0.f, wd = 0.f, unused = 0.f;
        ga.accumulate(acc0, ws0, unused_s, unused);
        if (CROSS) gb.accumulate(am, wm, ad, wd);
        for (int k = k0[0] + BA; k < k1[0]; k += BA) {        // long rows: the remaining entries
            GatherBatchG<BA, false, FS> g;
            g.load_entries(a.col[0], a.val[0], nullptr, k, k1[0]);
            g.load_rows(a.Xs);
            g.accumulate(acc0, ws0, unused_s, unused);
        }
        if (CROSS)
            for (int k = p0 + BP; k < p1; k += BP) {
                GatherBatchG<BP, true, FCC> g;
                g.load_entries(a.p_col, a.p_pm, a.p_pd, k, p1);
                g.load_rows(a.Xc);
                g.accumulate(am, wm, ad, wd);
            }
#pragma unroll
        for (int f = 0; f < FS; ++f) x1[2 * FS + f] = fmaf(acc0[f], bs.sc[f], ws0 * bs.sh[f]);
#pragma unroll
        for (int t = 1; t < NCSR; ++t) {
            float acc[FS];
#pragma unroll
            for (int f = 0; f < FS; ++f) acc[f] = 0.f;
            float ws = 0.f;
            for (int k = k0[t]; k < k1[t]; k += 4) {
                GatherBatchG<4, false, FS> g;
                g.load_entries(a.col[t], a.val[t], nullptr, k, k1[t]);
                g.load_rows(a.Xs);
                g.accumulate(acc, ws, unused_s, unused);
            }
#pragma unroll
            for (int f = 0; f < FS; ++f) x1[(2 + t) * FS + f] = fmaf(acc[f], bs.sc[f], ws * bs.sh[f]);
        }
        if (CROSS) {
#pragma unroll
            for (int f = 0; f < FCC; ++f) {
                x1[NS * FS + f] = fmaf(am[f], bc.sc[f], wm * bc.sh[f]);
                x1[NS * FS + FCC + f] = fmaf(ad[f], bc.sc[f], wd * bc.sh[f]);
            }
        }
        float out[4];
#pragma unroll
        for (int o = 0; o < FO; ++o) {
            float acc = bias[o];
#pragma unroll
            for (int c = 0; c < CIN; ++c) acc = fmaf(x1[c], W[o * CIN + c], acc);
            if (o >= a.relu_from) acc = fmaxf(acc, 0.f);
            out[o] = acc;
            s1[o] = fmaf(rw, acc, s1[o]);
            s2[o] = fmaf(rw * acc, acc, s2[o]);
        }
        if constexpr (FO == 4) {
            *reinterpret_cast<float4*>(a.Z + (size_t)rr * 4) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int o = 0; o < FO; ++o) a.Z[(size_t)rr * FO + o] = out[o];
        }
    }
    if (FO == 4 && a.acc_out) {
        const int lane = tid & 31, warp = tid >> 5;
        double st[8];
#pragma unroll
        for (int o = 0; o < 4; ++o) { st[o] = (double)s1[o]; st[4 + o] = (double)s2[o]; }
        warp_reduce_scatter<double, 8>(st);                   // lane l: total of value l >> 2
        if ((lane & 3) == 0) red[warp * 8 + (lane >> 2)] = st[0];
        __syncthreads();
        if (tid < 8) {
            double v = 0.0;
            for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * 8 + tid];
            accum_add(a.acc_out, 8, hgnn_ws_bins(8), tid, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward for the odd widths: gY is FG wide (4 with batch-norm + ReLU backward on the fly, or 1 / 2 for
// the readout: plain), the self input FS wide, the cross input FC wide (ignored when R_cross == 0).
// Same structure as bwd_row4_kernel (engine_row4.cuh), which stays the tuned path for 4 / 4 / 4.
// ---------------------------------------------------------------------------------------------
template <int FG>
struct GpreG {
    Gpre4 g4;                 // FG == 4
    const float* G;           // FG < 4: plain rows (no batch-norm, no ReLU: checked by the host)
    __device__ __forceinline__ RowV<FG> operator()(int row) const {
        RowV<FG> r;
        if constexpr (FG == 4) {
            const float4 t = g4(row);
            r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
        } else {
#pragma unroll
            for (int o = 0; o < FG; ++o) r.v[o] = __ldg(G + (size_t)row * FG + o);
        }
        return r;
    }
};

template <int NCSR, int FG, int FS, int FC>
__global__ void __launch_bounds__(R4_THREADS)
bwd_rowg_kernel(const Bwd4Args a) {
    constexpr int NT = 2 + NCSR;                               // self blocks
    constexpr int FCC = FC > 0 ? FC : 1;
    constexpr int NDS = NT * FG * FS, NDC = 2 * FG * FCC;      // dW accumulators of a self / cross thread
    constexpr int NDW = NDS > NDC ? NDS : NDC;
    static_assert(NDW <= 64, "dW accumulators must fit the 64-value flush");
    __shared__ float Ws[NT * FG * FS];                         // [t][o][f] = W[o][t*FS+f]
    __shared__ float Wc[2 * FG * FCC];                         // [t][o][f] = W[o][col0 + t*FC + f]
    __shared__ float red[(R4_THREADS / 32) * 64];
    __shared__ int flagged[R4_MAX_FLAGGED];
    __shared__ int n_flagged;
    __shared__ int rng_ids[R4_MAX_IDS];
    __shared__ float rng_sum[R4_MAX_IDS * FG];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_self = (int)blockIdx.x < a.ctas_self;
    pdl_launch_dependents();
    // ---- phase 0 (parameters only)
    if (tid == 0) n_flagged = 0;
    for (int i = tid; i < NT * FG * FS; i += R4_THREADS) {
        const int t = i / (FG * FS), o = (i / FS) % FG, f = i % FS;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Ws[i] = wrow[t * FS + f];
    }
    if (FC > 0 && a.R_cross > 0)
        for (int i = tid; i < 2 * FG * FCC; i += R4_THREADS) {
            const int t = i / (FG * FCC), o = (i / FCC) % FG, f = i % FCC;
            const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
            Wc[i] = wrow[a.col0_cross + t * FCC + f];
        }
    pdl_wait();
    // ---- coefficients of this side's BN + ReLU backward (FG == 4 only), the input's BN vectors (width 4 only)
    GpreG<FG> gp;
    gp.G = a.gY;
    if constexpr (FG == 4) {
        Gpre4& g4 = gp.g4;
        g4.relu_from = a.relu_from; g4.bn = a.has_bn != 0; g4.need_z = a.has_bn != 0 || a.relu_from < 4;
        g4.G = a.gY; g4.Z = a.Z;
        if (a.has_bn) {
            double tf[8], tb[8];
            warp_totals8(a.acc_f, tf);
            warp_totals8(a.acc_b, tb);
            const float w = a.bn_w[0];
            const double inv_n = 1.0 / (double)a.Rg;
            float c0[4], c1[4], c2[4];
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const double m = tf[f] * inv_n;
                const double var = fma(-m, m, tf[4 + f] * inv_n);
                const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
                const float k0 = w * r_;
                const float k2 = -k0 * (float)(tb[4 + f] * inv_n) * r_;
                c0[f] = k0;
                c2[f] = k2;
                c1[f] = -k0 * (float)(tb[f] * inv_n) - k2 * (float)m;
            }
            g4.c0 = make_float4(c0[0], c0[1], c0[2], c0[3]);
            g4.c1 = make_float4(c1[0], c1[1], c1[2], c1[3]);
            g4.c2 = make_float4(c2[0], c2[1], c2[2], c2[3]);
        } else {
            g4.c0 = make_float4(1.f, 1.f, 1.f, 1.f);
            g4.c1 = f4_zero();
            g4.c2 = f4_zero();
        }
    }
    BnV<FS> bxs;
    BnV<FCC> bxc;
    bxs.identity();
    bxc.identity();
    if (is_self) {
        if (FS == 4 && (a.bn_s.acc || a.bn_s.affine)) bxs.from4(bn4_from_ref(a.bn_s));
    } else {
        if (FC == 4 && (a.bn_c.acc || a.bn_c.affine)) bxc.from4(bn4_from_ref(a.bn_c));
    }
    __syncthreads();                                           // weights in shared memory

    float dw[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) dw[i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};

    if (is_self) {
        float* const gX = a.gXs;
        const bool stats = FS == 4 && a.acc_b_self != nullptr && gX != nullptr;
        // finishes one row from its T blocks: gX (+)= W^T T, dW += T (x) xn, statistics of the produced gradient
        auto finish = [&](int row, const float (&T)[NT][FG], bool add_to_existing, bool first_visit) {
            const RowV<FS> xr = ldrow<FS>(a.Xs, row);
            const float rw = a.roww_s ? __ldg(a.roww_s + row) : 1.f;      // weight of the row in every sum over rows
            float xn[FS], g[FS];
#pragma unroll
            for (int f = 0; f < FS; ++f) { xn[f] = rw * fmaf(xr.v[f], bxs.sc[f], bxs.sh[f]); g[f] = 0.f; }
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int o = 0; o < FG; ++o)
#pragma unroll
                    for (int f = 0; f < FS; ++f) {
                        g[f] = fmaf(T[t][o], Ws[(t * FG + o) * FS + f], g[f]);
                        dw[(t * FG + o) * FS + f] = fmaf(T[t][o], xn[f], dw[(t * FG + o) * FS + f]);
                    }
            if (first_visit) {
#pragma unroll
                for (int o = 0; o < FG; ++o) db[o] = fmaf(rw, T[0][o], db[o]);
            }
            if (gX) {
#pragma unroll
                for (int f = 0; f < FS; ++f) {
                    float v = g[f];
                    if (add_to_existing) v += __ldcg(gX + (size_t)row * FS + f);
                    gX[(size_t)row * FS + f] = v;
                }
                if (stats) {
#pragma unroll
                    for (int f = 0; f < FS; ++f) {
                        sg[f & 3] = fmaf(rw, g[f], sg[f & 3]);
                        sgx[f & 3] = fmaf(rw * g[f], (xr.v[f] - bxs.mu[f]) * bxs.rs[f], sgx[f & 3]);
                    }
                }
            }
        };
        for (int ridx = blockIdx.x * R4_THREADS + tid; ridx < a.R_self; ridx += a.ctas_self * R4_THREADS) {
            const int row = a.rowmap_s ? __ldg(a.rowmap_s + ridx) : ridx;
            if (a.roww_s && __ldg(a.roww_s + row) <= 0.f) continue;        // a skipped copy of a phantom line-graph row
            float T[NT][FG];
            const RowV<FG> t0 = gp(row);
            const float d = __ldg(a.diag + row);
#pragma unroll
            for (int o = 0; o < FG; ++o) { T[0][o] = t0.v[o]; T[1][o] = d * t0.v[o]; }
#pragma unroll
            for (int t = 0; t < NCSR; ++t) {
#pragma unroll
                for (int o = 0; o < FG; ++o) T[2 + t][o] = 0.f;
                const int k0 = __ldg(a.rowptr[t] + row), k1 = __ldg(a.rowptr[t] + row + 1);
                for (int k = k0; k < k1; k += 2) {
                    const bool on = k + 1 < k1;
                    const int c0 = __ldg(a.col[t] + k), c1 = on ? __ldg(a.col[t] + k + 1) : c0;
                    const float v0 = __ldg(a.val[t] + k), v1 = on ? __ldg(a.val[t] + k + 1) : 0.f;
                    const RowV<FG> g0 = gp(c0), g1 = gp(c1);
#pragma unroll
                    for (int o = 0; o < FG; ++o) T[2 + t][o] = fmaf(v1, g1.v[o], fmaf(v0, g0.v[o], T[2 + t][o]));
                }
            }
            if (a.rng_rowptr && __ldg(a.rng_rowptr + row + 1) > __ldg(a.rng_rowptr + row)) {
                const int slot = atomicAdd(&n_flagged, 1);
                if (slot < R4_MAX_FLAGGED) {
                    flagged[slot] = row;
                } else {            // list full: add the run-length part serially (correct, slow, rare)
                    for (int e = __ldg(a.rng_rowptr + row); e < __ldg(a.rng_rowptr + row + 1); ++e) {
                        const int id = __ldg(a.rng_id + e);
                        const float v = __ldg(a.rng_val + e);
                        for (int rr = __ldg(a.rng_lo + id); rr < __ldg(a.rng_hi + id); ++rr) {
                            const RowV<FG> gv = gp(rr);
#pragma unroll
                            for (int o = 0; o < FG; ++o) T[2][o] = fmaf(v, gv.v[o], T[2][o]);
                        }
                    }
                }
            }
            finish(row, T, a.acc_self != 0, true);
        }
        // ---- run-length parts of the flagged rows (as in bwd_row4_kernel): distinct range ids -> range sums,
        //      cooperatively -> every flagged row finished by its own thread
        __syncthreads();
        const int nf = min(n_flagged, R4_MAX_FLAGGED);
        if (nf > 0) {
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
                if (id < 0) break;                             // uniform
                const int lo = __ldg(a.rng_lo + id), hi = __ldg(a.rng_hi + id);
                float acc[FG];
#pragma unroll
                for (int o = 0; o < FG; ++o) acc[o] = 0.f;
                for (int rr = lo + tid; rr < hi; rr += 4 * R4_THREADS) {      // 4 rows in flight per thread
                    RowV<FG> gv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (rr + u * R4_THREADS < hi) {
                            gv[u] = gp(rr + u * R4_THREADS);
                        } else {
#pragma unroll
                            for (int o = 0; o < FG; ++o) gv[u].v[o] = 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int o = 0; o < FG; ++o) acc[o] += gv[u].v[o];
                }
#pragma unroll
                for (int o = 0; o < FG; ++o) {
                    acc[o] = warp_sum(acc[o]);
                    if (lane == 0) red[warp * FG + o] = acc[o];
                }
                __syncthreads();
                if (tid < FG) {
                    float v = 0.f;
                    for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * FG + tid];
                    rng_sum[j * FG + tid] = v;
                }
                __syncthreads();
            }
            if (tid < nf) {
                float T[NT][FG];
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int o = 0; o < FG; ++o) T[t][o] = 0.f;
                for (int e = e0; e < e1; ++e) {
                    const int id = __ldg(a.rng_id + e);
                    const float v = __ldg(a.rng_val + e);
                    int j = 0;
                    while (j < R4_MAX_IDS && rng_ids[j] != id) ++j;
                    if (j < R4_MAX_IDS) {
#pragma unroll
                        for (int o = 0; o < FG; ++o) T[2][o] = fmaf(v, rng_sum[j * FG + o], T[2][o]);
                    } else {        // more distinct ranges than table slots: serial sum (correct, slow, rare)
                        for (int rr = __ldg(a.rng_lo + id); rr < __ldg(a.rng_hi + id); ++rr) {
                            const RowV<FG> gv = gp(rr);
#pragma unroll
                            for (int o = 0; o < FG; ++o) T[2][o] = fmaf(v, gv.v[o], T[2][o]);
                        }
                    }
                }
                finish(my_row, T, true, false);                // the delta of everything that is linear in T[2]
            }
        }
    } else if (FC > 0) {
        float* const gX = a.gXc;
        const bool stats = FC == 4 && a.acc_b_cross != nullptr && gX != nullptr;
        const int ncta = gridDim.x - a.ctas_self;
        for (int ridx = (blockIdx.x - a.ctas_self) * R4_THREADS + tid; ridx < a.R_cross; ridx += ncta * R4_THREADS) {
            const int row = a.rowmap_c ? __ldg(a.rowmap_c + ridx) : ridx;
            const float rw = a.roww_c ? __ldg(a.roww_c + row) : 1.f;
            if (rw <= 0.f) continue;
            float T[2][FG];
#pragma unroll
            for (int o = 0; o < FG; ++o) { T[0][o] = 0.f; T[1][o] = 0.f; }
            const int k0 = __ldg(a.pt_rowptr + row), k1 = __ldg(a.pt_rowptr + row + 1);
            for (int k = k0; k < k1; k += 4) {
                int c[4];
                float vm[4], vd[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool on = k + j < k1;
                    c[j] = on ? __ldg(a.pt_col + k + j) : -1;
                    vm[j] = on ? __ldg(a.pt_pm + k + j) : 0.f;
                    vd[j] = on ? __ldg(a.pt_pd + k + j) : 0.f;
                }
                RowV<FG> gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (c[j] >= 0) {
                        gv[j] = gp(c[j]);
                    } else {
#pragma unroll
                        for (int o = 0; o < FG; ++o) gv[j].v[o] = 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int o = 0; o < FG; ++o) {
                        T[0][o] = fmaf(vm[j], gv[j].v[o], T[0][o]);
                        T[1][o] = fmaf(vd[j], gv[j].v[o], T[1][o]);
                    }
            }
            const RowV<FCC> xr = ldrow<FCC>(a.Xc, row);
            float xn[FCC], g[FCC];
#pragma unroll
            for (int f = 0; f < FCC; ++f) { xn[f] = rw * fmaf(xr.v[f], bxc.sc[f], bxc.sh[f]); g[f] = 0.f; }
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int o = 0; o < FG; ++o)
#pragma unroll
                    for (int f = 0; f < FCC; ++f) {
                        g[f] = fmaf(T[t][o], Wc[(t * FG + o) * FCC + f], g[f]);
                        dw[(t * FG + o) * FCC + f] = fmaf(T[t][o], xn[f], dw[(t * FG + o) * FCC + f]);
                    }
            if (gX) {
#pragma unroll
                for (int f = 0; f < FCC; ++f) {
                    float v = g[f];
                    if (a.acc_cross) v += gX[(size_t)row * FCC + f];
                    gX[(size_t)row * FCC + f] = v;
                }
                if (stats) {
#pragma unroll
                    for (int f = 0; f < FCC; ++f) {
                        sg[f & 3] = fmaf(rw, g[f], sg[f & 3]);
                        sgx[f & 3] = fmaf(rw * g[f], (xr.v[f] - bxc.mu[f]) * bxc.rs[f], sgx[f & 3]);
                    }
                }
            }
        }
    }
    // ---- flush: reduce-scatter butterflies -> per-warp rows in shared memory -> one fp64 atomic per value
    const int nvals = is_self ? NDS : NDC;
    const int fx = is_self ? FS : FCC;
    const int nbw = hgnn_ws_bins(FG * a.Cin);
    const int col_base = is_self ? 0 : a.col0_cross;
    __syncthreads();
    warp_reduce_scatter<float, 64>(dw);                        // lane l: totals of values 2l, 2l + 1
    red[warp * 64 + 2 * lane] = dw[0];
    red[warp * 64 + 2 * lane + 1] = dw[1];
    __syncthreads();
    if (a.dW_bins)
        for (int i = tid; i < nvals; i += R4_THREADS) {
            float v = 0.f;
            for (int w = 0; w < R4_THREADS / 32; ++w) v += red[w * 64 + i];
            const int t = i / (FG * fx), o = (i / fx) % FG, f = i % fx;
            accum_add(a.dW_bins, FG * a.Cin, nbw, o * a.Cin + col_base + t * fx + f, (double)v);
        }
    __syncthreads();
    float extra[16];
#pragma unroll
    for (int f = 0; f < 4; ++f) { extra[f] = db[f]; extra[4 + f] = sg[f]; extra[8 + f] = sgx[f]; extra[12 + f] = 0.f; }
    warp_reduce_scatter<float, 16>(extra);                     // lane l: total of value l >> 1
    if ((lane & 1) == 0) red[warp * 64 + (lane >> 1)] = extra[0];
    __syncthreads();
    if (tid < 12) {
        double v = 0.0;
        for (int w = 0; w < R4_THREADS / 32; ++w) v += (double)red[w * 64 + tid];
        if (tid < 4) {
            if (tid < FG && is_self && a.db_bins) accum_add(a.db_bins, FG, hgnn_ws_bins(FG), tid, v);
        } else {
            double* accb = is_self ? a.acc_b_self : a.acc_b_cross;
            float* gXp = is_self ? a.gXs : a.gXc;
            const bool wide4 = is_self ? FS == 4 : FC == 4;
            if (wide4 && accb && gXp) accum_add(accb, 8, hgnn_ws_bins(8), tid - 4, v);
        }
    }
}
