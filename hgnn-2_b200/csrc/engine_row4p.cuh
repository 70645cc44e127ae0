// engine_row4p.cuh -- the width-4 backward with its dependent load rounds taken off the critical path.
// Included by engine.cu inside namespace eng, after engine_row4.cuh (same Bwd4Args, same arithmetic per row).
//
// profiles/cta_phases.py on bwd_row4_kernel (C2, eager launches, us per CTA, median): 1.95 from the CTA start to
// "coefficient vectors in shared memory", 4.0-5.7 in the row loop, 1.0-1.5 in the flush - of a 8.8 (node side) /
// 11.4 (edge side) us kernel.  The row loop is a chain of dependent L2 round trips of ~0.6 us each:
//     gPre(row) -> row pointers -> entries -> gPre(col)      per batch of GB / CB entries again: entries -> gPre(col)
// (a node row of Pm^T / Pd^T has 8.3 entries = 3 batches of 4 = 6 rounds; an A^T row 5 entries = 3 batches of 2), and
// every thread of the heavier part walks two rows.  Nothing in that chain but the gPre loads depends on the producer
// kernel.  So here:
//   * before griddepcontrol.wait (under the producer's tail with PDL): the structure of the thread's first TWO rows -
//     row list entry, weight, diagonal, row pointers AND the first batch of (col, val) entries of each;
//   * right after the wait, BEFORE the batch-norm / gPre coefficient vectors are derived: every producer-written
//     load of the first row - its own (g, z), x and the (g, z) rows of the first batch -
//     so the coefficient round (accumulator load, fp64 sums, CTA barrier) runs under them instead of in front;
//   * inside a row, the entries of batch b + 1 are requested before the (g, z) rows of batch b are consumed: one
//     round per batch instead of two;
//   * the structure of row i + 1 is requested before row i is computed;
//   * one flush: dW, dbias and the batch-norm sums go through ONE reduce-scatter + ONE CTA barrier (was four).
// Rows with run-length parts (the uncollapsed line graph) and the gather-only variant stay on bwd_row4_kernel.
#pragma once


template <int B, bool TWO>
struct Ent4 {
    int c[B];
    float v[B], v2[TWO ? B : 1];
    __device__ __forceinline__ void off() {
#pragma unroll
        for (int j = 0; j < B; ++j) { c[j] = -1; v[j] = 0.f; if (TWO) v2[j] = 0.f; }
    }
    __device__ __forceinline__ void load(const int* __restrict__ col, const float* __restrict__ val,
                                         const float* __restrict__ val2, int k, int k1) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const bool on = k + j < k1;
            c[j] = on ? __ldg(col + k + j) : -1;
            v[j] = on ? __ldg(val + k + j) : 0.f;
            if (TWO) v2[j] = on ? __ldg(val2 + k + j) : 0.f;
        }
    }
};

struct Row4S { int row; float rw, d; int k0, k1; };

// gPre split into its loads and its arithmetic (Gpre4::operator() = fin(raw))
__device__ __forceinline__ void gpre_raw(const Gpre4& gp, int row, float4& g, float4& z) {
    g = ld4(gp.G + (size_t)row * 4);
    z = gp.need_z ? ld4(gp.Z + (size_t)row * 4) : f4_zero();
}
// the coefficient vectors c0, c1, c2 are read from shared memory (cv[0..12)) at the point of use: held in registers
// they would be 12 more values alive across every load round
__device__ __forceinline__ float4 gpre_fin(const Gpre4& gp, const float* __restrict__ cv, float4 g, float4 z) {
    if (!gp.need_z) return g;
    if (gp.bn) {
        const float4 c0 = *reinterpret_cast<const float4*>(cv), c1 = *reinterpret_cast<const float4*>(cv + 4);
        const float4 c2 = *reinterpret_cast<const float4*>(cv + 8);
        g = make_float4(fmaf(c2.x, z.x, fmaf(c0.x, g.x, c1.x)), fmaf(c2.y, z.y, fmaf(c0.y, g.y, c1.y)),
                        fmaf(c2.z, z.z, fmaf(c0.z, g.z, c1.z)), fmaf(c2.w, z.w, fmaf(c0.w, g.w, c1.w)));
    }
    if (0 >= gp.relu_from && !(z.x > 0.f)) g.x = 0.f;
    if (1 >= gp.relu_from && !(z.y > 0.f)) g.y = 0.f;
    if (2 >= gp.relu_from && !(z.z > 0.f)) g.z = 0.f;
    if (3 >= gp.relu_from && !(z.w > 0.f)) g.w = 0.f;
    return g;
}

// One part of the launch (CTA-uniform): SELF = transposed [IDENT, DIAG, CSR] on the rows of the side's own input,
// !SELF = the Pm^T / Pd^T pattern on the rows of the cross input.  B entries per gather batch.
// R4P_RMW_LOAD: accumulate into gX with load + add + store (the load goes out with the row's other loads) instead of
// red.global.add.v4.f32 (no load, 4 registers fewer - but the reductions of 80 k rows drain for ~1 us after the last CTA,
// which delays the release of the dependent launch: profiles/logs/step_timeline_r3e.log)
template <int B, bool SELF, bool R4P_RMW_LOAD>
__device__ __forceinline__ void bwd4p_part(const Bwd4Args& a, const float* __restrict__ W, float* __restrict__ red,
                                           float* __restrict__ cv) {
    constexpr int NT = SELF ? 3 : 2;
    constexpr bool TWO = !SELF;
    constexpr bool PRE2 = SELF || B <= 2;      // entries of the row after the current one requested a row ahead (3 B registers
                                               // with two value arrays: with B = 4 they spill under the 128-register bound)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = SELF ? a.R_self : a.R_cross;
    const int* __restrict__ rowmap = SELF ? a.rowmap_s : a.rowmap_c;
    const float* __restrict__ roww = SELF ? a.roww_s : a.roww_c;
    const int* __restrict__ rowptr = SELF ? a.rowptr[0] : a.pt_rowptr;
    const int* __restrict__ col = SELF ? a.col[0] : a.pt_col;
    const float* __restrict__ val = SELF ? a.val[0] : a.pt_pm;
    const float* __restrict__ val2 = SELF ? nullptr : a.pt_pd;
    const float* __restrict__ X = SELF ? a.Xs : a.Xc;
    float* const gX = SELF ? a.gXs : a.gXc;
    const bool accum = (SELF ? a.acc_self : a.acc_cross) != 0;
    double* const accb = SELF ? a.acc_b_self : a.acc_b_cross;
    const bool stats = accb != nullptr && gX != nullptr;
    const int cta = SELF ? (int)blockIdx.x : (int)blockIdx.x - a.ctas_self;
    const int ncta = SELF ? a.ctas_self : (int)gridDim.x - a.ctas_self;
    const int stride = ncta * R4_THREADS;

    auto load_struct = [&](int idx) {
        Row4S s;
        s.row = -1; s.rw = 0.f; s.d = 0.f; s.k0 = 0; s.k1 = 0;
        if (idx < R) {
            s.row = rowmap ? __ldg(rowmap + idx) : idx;
            s.rw = roww ? __ldg(roww + s.row) : 1.f;
            if (SELF) s.d = __ldg(a.diag + s.row);
            s.k0 = __ldg(rowptr + s.row);
            s.k1 = __ldg(rowptr + s.row + 1);
            if (s.rw <= 0.f) s.k1 = s.k0;          // a skipped copy of a phantom line-graph row: no gather
        }
        return s;
    };
    auto load_entries = [&](const Row4S& s, Ent4<B, TWO>& e) {
        if (s.row >= 0) e.load(col, val, val2, s.k0, s.k1);
        else e.off();
    };

    // ---- structure of the first two rows: nothing the producer writes
    int nidx = cta * R4_THREADS + tid;
    Row4S cur = load_struct(nidx); nidx += stride;
    Row4S nxt = load_struct(nidx); nidx += stride;
    Ent4<B, TWO> curE, nxtE;
    load_entries(cur, curE);
    if (PRE2) load_entries(nxt, nxtE);

    pdl_wait();
    ktrace_waited(a.trace_slot);

    // ---- every producer-written load of the first row, before the coefficient round
    Gpre4 gp;
    gp.relu_from = a.relu_from; gp.bn = a.has_bn != 0; gp.need_z = a.has_bn != 0 || a.relu_from < 4;
    gp.G = a.gY; gp.Z = a.Z;
    float4 own_g = f4_zero(), own_z = f4_zero(), xr = f4_zero(), old = f4_zero();
    float4 dg[B], dz[B];
    auto issue = [&](const Row4S& s, const Ent4<B, TWO>& e) {
        if (SELF) gpre_raw(gp, s.row, own_g, own_z);
        xr = ld4(X + (size_t)s.row * 4);
        if (R4P_RMW_LOAD && gX && accum) old = __ldcg(reinterpret_cast<const float4*>(gX + (size_t)s.row * 4));
#pragma unroll
        for (int j = 0; j < B; ++j) {
            if (e.c[j] >= 0) gpre_raw(gp, e.c[j], dg[j], dz[j]);
            else { dg[j] = f4_zero(); dz[j] = f4_zero(); }
        }
    };
    if (cur.row >= 0) issue(cur, curE);

    // ---- coefficients of this side's BN + ReLU backward (warp 0) and the input's BN vectors (warp 1), once per CTA
    if (warp == 0) {
        float c0 = 1.f, c1 = 0.f, c2 = 0.f;
        if (a.has_bn) {     // one feature per lane: (sum z, sum z^2) and (sum g, sum g xhat) of this side's output
            double vf = hgnn_bins8_lane(a.acc_f), vb = hgnn_bins8_lane(a.acc_b);
            const float w = a.bn_w[0];
            const int f = lane & 3;
            vf += __shfl_xor_sync(0xffffffffu, vf, 8); vb += __shfl_xor_sync(0xffffffffu, vb, 8);
            vf += __shfl_xor_sync(0xffffffffu, vf, 16); vb += __shfl_xor_sync(0xffffffffu, vb, 16);
            const double sum = __shfl_sync(0xffffffffu, vf, f), sq = __shfl_sync(0xffffffffu, vf, 4 + f);
            const double sgt = __shfl_sync(0xffffffffu, vb, f), sgx = __shfl_sync(0xffffffffu, vb, 4 + f);
            const double inv_n = a.inv_Rg;
            const double m = sum * inv_n;
            const double var = fma(-m, m, sq * inv_n);
            const float r_ = 1.0f / sqrtf(fmaxf((float)var, 0.f) + (float)ENG_BN_EPS);
            const float k0 = w * r_;
            const float k2 = -k0 * (float)(sgx * inv_n) * r_;
            c0 = k0;
            c2 = k2;
            c1 = -k0 * (float)(sgt * inv_n) - k2 * (float)m;
        }
        if (lane < 4) { cv[lane] = c0; cv[4 + lane] = c1; cv[8 + lane] = c2; }
    } else if (warp == 1) {
        const BnRef& r = SELF ? a.bn_s : a.bn_c;
        if (!r.affine && r.acc) {
            float sc, sh, mu, rs;
            bn4_lane(hgnn_bins8_lane(r.acc), r.w[0], r.b[0], r.inv_n, sc, sh, mu, rs);
            if (lane < 4) { cv[12 + lane] = sc; cv[16 + lane] = sh; cv[20 + lane] = mu; cv[24 + lane] = rs; }
        } else {
            const Bn4 bx = bn4_from_ref(r);
            if (lane == 0) {
                *reinterpret_cast<float4*>(cv + 12) = bx.sc; *reinterpret_cast<float4*>(cv + 16) = bx.sh;
                *reinterpret_cast<float4*>(cv + 20) = bx.mu; *reinterpret_cast<float4*>(cv + 24) = bx.rs;
            }
        }
    }
    __syncthreads();                                   // weights and coefficient vectors in shared memory
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_phase[blockIdx.x * 3 + 2] = global_ns();

    float dw[NT * 16];
#pragma unroll
    for (int i = 0; i < NT * 16; ++i) dw[i] = 0.f;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgx[4] = {0.f, 0.f, 0.f, 0.f};

    // arithmetic of one row whose loads are in flight (own_g / own_z / xr / dg / dz)
    auto compute = [&](const Row4S& cur, const Ent4<B, TWO>& curE) {
        if (cur.rw > 0.f) {
            float4 T[3];
            if (SELF) {
                T[0] = gpre_fin(gp, cv, own_g, own_z);
                T[1] = make_float4(cur.d * T[0].x, cur.d * T[0].y, cur.d * T[0].z, cur.d * T[0].w);
            }
            float4 Ta = f4_zero(), Tb = f4_zero();
            {   // gather: entries of batch b + 1 requested before the rows of batch b are consumed
                Ent4<B, TWO> e = curE;
                for (int k = cur.k0 + B;; k += B) {
                    const bool more = k < cur.k1;
                    Ent4<B, TWO> en;
                    if (more) en.load(col, val, val2, k, cur.k1);
#pragma unroll
                    for (int j = 0; j < B; ++j) {
                        const float4 gv = gpre_fin(gp, cv, dg[j], dz[j]);
                        Ta = f4_fma(e.v[j], gv, Ta);
                        if (TWO) Tb = f4_fma(e.v2[j], gv, Tb);
                    }
                    if (!more) break;
#pragma unroll
                    for (int j = 0; j < B; ++j) {
                        if (en.c[j] >= 0) gpre_raw(gp, en.c[j], dg[j], dz[j]);
                        else { dg[j] = f4_zero(); dz[j] = f4_zero(); }
                    }
                    e = en;
                }
            }
            if (SELF) T[2] = Ta; else { T[0] = Ta; T[1] = Tb; }
            const float rw = cur.rw;
            const float4 xn1 = f4_affine(xr, *reinterpret_cast<const float4*>(cv + 12), *reinterpret_cast<const float4*>(cv + 16));
            const float4 xn = make_float4(rw * xn1.x, rw * xn1.y, rw * xn1.z, rw * xn1.w);   // weight of the row in dW
            float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float Tv[4] = {T[t].x, T[t].y, T[t].z, T[t].w};
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float4 w = *reinterpret_cast<const float4*>(W + (t * 4 + o) * 4);
                    g[0] = fmaf(Tv[o], w.x, g[0]); g[1] = fmaf(Tv[o], w.y, g[1]);
                    g[2] = fmaf(Tv[o], w.z, g[2]); g[3] = fmaf(Tv[o], w.w, g[3]);
                    dw[(t * 4 + o) * 4 + 0] = fmaf(Tv[o], xn.x, dw[(t * 4 + o) * 4 + 0]);
                    dw[(t * 4 + o) * 4 + 1] = fmaf(Tv[o], xn.y, dw[(t * 4 + o) * 4 + 1]);
                    dw[(t * 4 + o) * 4 + 2] = fmaf(Tv[o], xn.z, dw[(t * 4 + o) * 4 + 2]);
                    dw[(t * 4 + o) * 4 + 3] = fmaf(Tv[o], xn.w, dw[(t * 4 + o) * 4 + 3]);
                }
            }
            if (SELF) { db[0] = fmaf(rw, T[0].x, db[0]); db[1] = fmaf(rw, T[0].y, db[1]); db[2] = fmaf(rw, T[0].z, db[2]); db[3] = fmaf(rw, T[0].w, db[3]); }
            if (gX) {
                // every row has exactly one writer per launch: the vector reduction is the same sum as load + add + store,
                // without the load (red.global.add.v4.f32)
                float4 o4 = make_float4(g[0], g[1], g[2], g[3]);
                if (R4P_RMW_LOAD) {
                    if (accum) { o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w; }
                    *reinterpret_cast<float4*>(gX + (size_t)cur.row * 4) = o4;
                } else {
                    if (accum) atomicAdd(reinterpret_cast<float4*>(gX + (size_t)cur.row * 4), o4);
                    else *reinterpret_cast<float4*>(gX + (size_t)cur.row * 4) = o4;
                }
                if (stats) {
                    const float4 mu = *reinterpret_cast<const float4*>(cv + 20), rs = *reinterpret_cast<const float4*>(cv + 24);
                    const float xh[4] = {(xr.x - mu.x) * rs.x, (xr.y - mu.y) * rs.y, (xr.z - mu.z) * rs.z, (xr.w - mu.w) * rs.w};
#pragma unroll
                    for (int f = 0; f < 4; ++f) { sg[f] = fmaf(rw, g[f], sg[f]); sgx[f] = fmaf(rw * g[f], xh[f], sgx[f]); }
                }
            }
        }
    };
    // The first row is peeled: its loads went out before the coefficient round, and the accumulators start from its
    // products.  Later rows request their loads at the top of the iteration (nothing in flight across the back edge:
    // the accumulators already take 60 registers) with the structure of the row after them behind it.
    if (cur.row >= 0) {
        compute(cur, curE);
        cur = nxt;
        if (PRE2) curE = nxtE; else load_entries(cur, curE);
        while (cur.row >= 0) {
            issue(cur, curE);
            nxt = load_struct(nidx);
            nidx += stride;
            if (PRE2) load_entries(nxt, nxtE);
            compute(cur, curE);
            cur = nxt;
            if (PRE2) curE = nxtE; else load_entries(cur, curE);
        }
    }
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_phase[blockIdx.x * 3] = g_cta_phase[blockIdx.x * 3 + 1] = global_ns();

    // ---- one flush: [dW (NT * 16) | dbias (4, self) | sum g (4) | sum g xhat (4)] -> reduce-scatter -> shared memory
    //      -> one barrier -> one fp64 atomic per value
    float pad[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) pad[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NT * 16; ++i) pad[i] = dw[i];
#pragma unroll
    for (int f = 0; f < 4; ++f) { pad[48 + f] = db[f]; pad[52 + f] = sg[f]; pad[56 + f] = sgx[f]; }
    warp_reduce_scatter<float, 64>(pad);                 // lane l: totals of values 2l, 2l + 1
    red[warp * 64 + 2 * lane] = pad[0];
    red[warp * 64 + 2 * lane + 1] = pad[1];
    __syncthreads();
    if (tid < 60) {
        float vf = 0.f;
#pragma unroll
        for (int w = 0; w < R4_THREADS / 32; ++w) vf += red[w * 64 + tid];
        const double v = (double)vf;
        if (tid < 48) {
            if (tid < NT * 16 && a.dW_bins) {
                const int t = tid >> 4, o = (tid >> 2) & 3, f = tid & 3;
                const int col_base = SELF ? 0 : a.col0_cross;
                accum_add(a.dW_bins, 4 * a.Cin, hgnn_ws_bins(4 * a.Cin), o * a.Cin + col_base + t * 4 + f, v);
            }
        } else if (tid < 52) {
            if (SELF && a.db_bins) accum_add(a.db_bins, 4, hgnn_ws_bins(4), tid - 48, v);
        } else if (stats) {
            accum_add(accb, 8, hgnn_ws_bins(8), tid - 52, v);        // sum g (0..3), sum g xhat (4..7)
        }
    }
}

template <int GB, int CB, bool RMW>
__global__ void __launch_bounds__(R4_THREADS, R4_BWD_MIN_CTAS)
bwd_row4p_kernel(const Bwd4Args a) {
    __shared__ __align__(16) float Ws[3 * 4 * 4];         // [t][o][f] = W[o][t*4+f]
    __shared__ __align__(16) float Wc[2 * 4 * 4];         // [t][o][f] = W[o][col0 + t*4 + f]
    __shared__ float red[(R4_THREADS / 32) * 64];
    __shared__ __align__(16) float cv[28];                // c0, c1, c2 | scale, shift, mean, 1/std
    const int tid = threadIdx.x;
    const bool is_self = (int)blockIdx.x < a.ctas_self;
    pdl_launch_dependents();
    ktrace_start(a.trace_slot);
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) {
        g_cta_times[blockIdx.x * 3] = global_ns();
        g_cta_times[blockIdx.x * 3 + 2] = is_self ? 1 : 0;
    }
    // ---- parameters only
    if (is_self) {
        for (int i = tid; i < 48; i += R4_THREADS) {
            const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
            const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
            Ws[i] = wrow[t * 4 + f];
        }
        bwd4p_part<GB, true, RMW>(a, Ws, red, cv);
    } else {
        for (int i = tid; i < 32; i += R4_THREADS) {
            const int t = i >> 4, o = (i >> 2) & 3, f = i & 3;
            const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
            Wc[i] = wrow[a.col0_cross + t * 4 + f];
        }
        bwd4p_part<CB, false, RMW>(a, Wc, red, cv);
    }
    if ((a.ablate & 8) && tid == 0 && blockIdx.x < 2048) g_cta_times[blockIdx.x * 3 + 1] = global_ns();
    ktrace_end(a.trace_slot);
}
