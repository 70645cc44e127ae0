// engine_quad.cuh -- width-4 engine kernels with SEVERAL LANES PER ROW (h = 2, one CSR operator: J = 1).
// Included by engine.cu inside namespace eng, after engine_row4.cuh (shares its argument blocks and helpers).
//
// Why: the thread-per-row kernels of engine_row4.cuh are bound by the instruction stream of ONE thread - a row is
// ~400 (forward) / ~500+ (backward) dependent-ish instructions, 125-255 registers keep 8-16 warps on an SM, ncu shows
// issue slots 20-27 % busy and 10-21 % of the warp slots active (profiles/prof_row4_r2q_raw.csv); the persistent
// kernels of mega.cu confirmed it (profiles/README.md, round 2, item 2).  Here LPR = 2 or 4 adjacent lanes share a row:
//   * the gathers are split by entry (lane q takes entries q, q + LPR, ...), so a row's loads go out LPR times wider;
//   * the forward sums PARTIAL OUTPUTS (4 values, 2 shuffle steps) instead of the 15 partial block sums - every block
//     of x1 is linear in its gather sums, so each lane pushes its share through the weights first;
//   * the backward is parallel over the input feature f: after the (cheap) reduction of the gathered T block every
//     lane needs all of T but produces only ITS columns of gX, dW (12 or 24 accumulators instead of 48), the
//     batch-norm sums and dbias - registers drop to ~64 and 8 CTAs of 128 threads fit an SM.
// Same arithmetic as engine_row4.cuh up to the order of fp32 sums inside a row.
//
// MEASURED (round 2, C2): parity green, but SLOWER than one thread per row - forward node side 11.4 vs 7.6 us, edge
// side 16.0 vs 7.2 us (4 lanes), 10.6 us both (2 lanes).  ncu: 4.2 M warp instructions against 1.6 M (every lane
// repeats structure loads, addressing and the batch-norm prologue; 8 rows advance per warp instruction instead of 32)
// at 37 % issue utilisation against 21 %.  These kernels are bound by instructions executed at a low issue rate, not
// by the length of one thread's chain: more lanes per row buy issue rate slower than they add instructions.  The
// forward stays as an opt-in experiment (HGNN_B200_QUAD=1); the backward counterpart was not written.
#pragma once

template <int B, bool TWO, int LPR>
struct QuadBatch {      // B entries per lane: entry j of lane q is k + q + LPR * j
    int c[B];
    float v[B], v2[TWO ? B : 1];
    float4 x[B];
    __device__ __forceinline__ void load_entries(const int* __restrict__ col, const float* __restrict__ val,
                                                 const float* __restrict__ val2, int k, int k1, int q) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const int e = k + q + LPR * j;
            const bool on = e < k1;
            c[j] = on ? __ldg(col + e) : -1;
            v[j] = on ? __ldg(val + e) : 0.f;
            if (TWO) v2[j] = on ? __ldg(val2 + e) : 0.f;
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

// sum over the LPR lanes of a row group (xor butterflies: every lane ends with the total)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int m = 1; m < LPR; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
// sum over the lanes of a warp that hold the same position q in their group (lanes q, q + LPR, ...)
template <int LPR, typename T>
__device__ __forceinline__ T column_sum(T v) {
#pragma unroll
    for (int m = LPR; m < 32; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

#define QD_THREADS 128

// ---------------------------------------------------------------------------------------------
// forward: out = [relu](W [x, d x, A x | Pm xc, Pd xc] + b), raw, + weighted (sum z, sum z^2)
// JA / JP: entries per lane and batch of the CSR operator / the incidence pair
// ---------------------------------------------------------------------------------------------
template <bool CROSS, int LPR, int JA, int JP>
__global__ void __launch_bounds__(QD_THREADS, 8)
fwd_quad_kernel(const Fwd4Args a) {
    constexpr int NB = 3 + (CROSS ? 2 : 0);
    constexpr int FPL = 4 / LPR;                         // outputs owned by a lane (store + statistics)
    constexpr int RPC = QD_THREADS / LPR;                // rows per CTA and iteration
    __shared__ __align__(16) float W[4 * NB * 4];        // [o][Cin]
    __shared__ __align__(16) float bias[4];
    __shared__ double red[(QD_THREADS / 32) * 8];
    const int tid = threadIdx.x, q = tid & (LPR - 1);
    pdl_launch_dependents();
    // ---- phase 0: parameters and graph structure only (overlaps the producer's tail under PDL)
    for (int i = tid; i < 4 * NB * 4; i += QD_THREADS) {
        const int o = i / (NB * 4), c = i - o * (NB * 4);
        W[i] = (o < a.Ha) ? a.Wa[(size_t)o * a.Cin + c] : a.Wb[(size_t)(o - a.Ha) * a.Cin + c];
    }
    if (tid < 4) bias[tid] = (tid < a.Ha) ? (a.ba ? a.ba[tid] : 0.f) : (a.bb ? a.bb[tid - a.Ha] : 0.f);
    Bn4Loader ls, lc;
    ls.issue_params(a.bn_s);
    if (CROSS) lc.issue_params(a.bn_c);
    const int stride = gridDim.x * RPC;
    int ridx = blockIdx.x * RPC + tid / LPR;
    int rr = 0, k0 = 0, k1 = 0, p0 = 0, p1 = 0;
    float d = 0.f, rw = 0.f;
    QuadBatch<JA, false, LPR> ga;
    QuadBatch<JP, true, LPR> gb;
    auto load_structure = [&](int idx) {
        rr = a.rowmap ? __ldg(a.rowmap + idx) : idx;
        d = __ldg(a.diag + rr);
        k0 = __ldg(a.rowptr[0] + rr);
        k1 = __ldg(a.rowptr[0] + rr + 1);
        if (CROSS) { p0 = __ldg(a.p_rowptr + rr); p1 = __ldg(a.p_rowptr + rr + 1); }
        rw = a.roww ? __ldg(a.roww + rr) : 1.f;
        if (rw <= 0.f) { k1 = k0; p1 = p0; }
        ga.load_entries(a.col[0], a.val[0], nullptr, k0, k1, q);
        if (CROSS) gb.load_entries(a.p_col, a.p_pm, a.p_pd, p0, p1, q);
    };
    auto no_row = [&]() {                                // a group without a row still takes part in the shuffles
        k0 = k1 = p0 = p1 = 0;
        rw = 0.f;
        ga.load_entries(a.col[0], a.val[0], nullptr, 0, 0, q);
        if (CROSS) gb.load_entries(a.p_col, a.p_pm, a.p_pd, 0, 0, q);
    };
    if (ridx < a.R) load_structure(ridx);
    else no_row();
    // ---- phase 1: everything the producer wrote
    pdl_wait();
    ls.issue_acc(a.bn_s);
    if (CROSS) lc.issue_acc(a.bn_c);
    float4 xs_raw = f4_zero();
    if (ridx < a.R) xs_raw = ld4(a.Xs + (size_t)rr * 4);
    ga.load_rows(a.Xs);                                  // a group without a row reads nothing (all columns -1)
    if (CROSS) gb.load_rows(a.Xc);
    const Bn4 bs4 = ls.resolve(a.bn_s);
    Bn4 bc4 = bs4;
    if (CROSS) bc4 = lc.resolve(a.bn_c);
    __syncthreads();                                   // weights in shared memory
    const float4 sc_s = bs4.sc, sh_s = bs4.sh, sc_c = bc4.sc, sh_c = bc4.sh;
    float s1[FPL], s2[FPL];
#pragma unroll
    for (int f = 0; f < FPL; ++f) { s1[f] = 0.f; s2[f] = 0.f; }

    // whole warps iterate together (the group sums are full-warp shuffles): a group past the end, or on a skipped copy
    // of a phantom line-graph row, computes on zeros and neither stores nor counts
    for (bool first = true; __any_sync(0xffffffffu, ridx < a.R); ridx += stride, first = false) {
        if (!first) {
            if (ridx < a.R) {
                load_structure(ridx);
                xs_raw = ld4(a.Xs + (size_t)rr * 4);
            } else {
                no_row();
            }
            ga.load_rows(a.Xs);
            if (CROSS) gb.load_rows(a.Xc);
        }
        const bool active = ridx < a.R && rw > 0.f;
        float4 acc0 = f4_zero(), am = f4_zero(), ad = f4_zero(), u4 = f4_zero();
        float ws0 = 0.f, wm = 0.f, wd = 0.f, u = 0.f;
        ga.accumulate(acc0, ws0, u4, u);
        if (CROSS) gb.accumulate(am, wm, ad, wd);
        for (int k = k0 + JA * LPR; k < k1; k += JA * LPR) {      // long rows: the remaining entries, batch by batch
            QuadBatch<JA, false, LPR> g;
            g.load_entries(a.col[0], a.val[0], nullptr, k, k1, q);
            g.load_rows(a.Xs);
            g.accumulate(acc0, ws0, u4, u);
        }
        if (CROSS)
            for (int k = p0 + JP * LPR; k < p1; k += JP * LPR) {
                QuadBatch<JP, true, LPR> g;
                g.load_entries(a.p_col, a.p_pm, a.p_pd, k, p1, q);
                g.load_rows(a.Xc);
                g.accumulate(am, wm, ad, wd);
            }
        // this lane's share of the gathered blocks, normalised: sum val*(s z + t) = s (sum val z) + t (sum val)
        float4 xp[NB - 2];
        xp[0] = make_float4(fmaf(acc0.x, sc_s.x, ws0 * sh_s.x), fmaf(acc0.y, sc_s.y, ws0 * sh_s.y),
                            fmaf(acc0.z, sc_s.z, ws0 * sh_s.z), fmaf(acc0.w, sc_s.w, ws0 * sh_s.w));
        if (CROSS) {
            xp[1] = make_float4(fmaf(am.x, sc_c.x, wm * sh_c.x), fmaf(am.y, sc_c.y, wm * sh_c.y),
                                fmaf(am.z, sc_c.z, wm * sh_c.z), fmaf(am.w, sc_c.w, wm * sh_c.w));
            xp[2] = make_float4(fmaf(ad.x, sc_c.x, wd * sh_c.x), fmaf(ad.y, sc_c.y, wd * sh_c.y),
                                fmaf(ad.z, sc_c.z, wd * sh_c.z), fmaf(ad.w, sc_c.w, wd * sh_c.w));
        }
        const float4 xs = f4_affine(xs_raw, sc_s, sh_s);
        const float4 xd = make_float4(d * xs.x, d * xs.y, d * xs.z, d * xs.w);
        // partial outputs: the gathered blocks for all 4 outputs, the own-row blocks + bias for the outputs this lane owns
        float out[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = 0.f;
#pragma unroll
            for (int b = 0; b < NB - 2; ++b) acc += f4_dot(xp[b], *reinterpret_cast<const float4*>(W + (o * NB + 2 + b) * 4));
            if (o / FPL == q)
                acc += bias[o] + f4_dot(xs, *reinterpret_cast<const float4*>(W + (o * NB) * 4)) +
                       f4_dot(xd, *reinterpret_cast<const float4*>(W + (o * NB + 1) * 4));
            out[o] = group_sum<LPR>(acc);
            if (o >= a.relu_from) out[o] = fmaxf(out[o], 0.f);
        }
        if (!active) continue;
        // store and statistics of the outputs this lane owns
        if (LPR == 4) {
            const float mine = q == 0 ? out[0] : q == 1 ? out[1] : q == 2 ? out[2] : out[3];
            a.Z[(size_t)rr * 4 + q] = mine;
            s1[0] = fmaf(rw, mine, s1[0]);
            s2[0] = fmaf(rw * mine, mine, s2[0]);
        } else if (LPR == 2) {
            const float m0 = q == 0 ? out[0] : out[2], m1 = q == 0 ? out[1] : out[3];
            *reinterpret_cast<float2*>(a.Z + (size_t)rr * 4 + 2 * q) = make_float2(m0, m1);
            s1[0] = fmaf(rw, m0, s1[0]); s2[0] = fmaf(rw * m0, m0, s2[0]);
            s1[FPL - 1] = fmaf(rw, m1, s1[FPL - 1]); s2[FPL - 1] = fmaf(rw * m1, m1, s2[FPL - 1]);
        } else {
            *reinterpret_cast<float4*>(a.Z + (size_t)rr * 4) = make_float4(out[0], out[1], out[2], out[3]);
#pragma unroll
            for (int o = 0; o < FPL; ++o) { s1[o] = fmaf(rw, out[o], s1[o]); s2[o] = fmaf(rw * out[o], out[o], s2[o]); }
        }
    }
    if (a.acc_out) {   // per feature: lanes with the same q -> one value per warp -> 8 fp64 atomics per CTA
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int f = 0; f < FPL; ++f) {
            const double t1 = column_sum<LPR>((double)s1[f]), t2 = column_sum<LPR>((double)s2[f]);
            if (lane < LPR) { red[warp * 8 + lane * FPL + f] = t1; red[warp * 8 + 4 + lane * FPL + f] = t2; }
        }
        __syncthreads();
        if (tid < 8) {
            double v = 0.0;
            for (int w = 0; w < QD_THREADS / 32; ++w) v += red[w * 8 + tid];
            accum_add(a.acc_out, 8, hgnn_ws_bins(8), tid, v);
        }
    }
}
