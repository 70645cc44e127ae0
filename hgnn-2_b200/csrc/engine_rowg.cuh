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
    float d = 0.f;
    int k0[NCSR], k1[NCSR], p0 = 0, p1 = 0;
    GatherBatchG<BA, false, FS> ga;
    GatherBatchG<BP, true, FCC> gb;
    auto load_structure = [&](int r) {
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
        xs_raw = ldrow<FS>(a.Xs, row);
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
            xs_raw = ldrow<FS>(a.Xs, row);
            ga.load_rows(a.Xs);
            if (CROSS) gb.load_rows(a.Xc);
        }
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
            s1[o] += acc;
            s2[o] = fmaf(acc, acc, s2[o]);
        }
        if constexpr (FO == 4) {
            *reinterpret_cast<float4*>(a.Z + (size_t)row * 4) = make_float4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int o = 0; o < FO; ++o) a.Z[(size_t)row * FO + o] = out[o];
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
