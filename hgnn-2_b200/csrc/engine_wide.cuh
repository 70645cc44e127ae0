// engine_wide.cuh -- tile kernels for WIDE states (16 | Fout, the h >= 16 regime), included by engine.cu
// inside namespace eng.
//
// At these widths a launch moves hundreds of MB and the per-row linear (320 x 64 at h = 32) is a real dense
// contraction, so the two things that bound the generic tile kernels are different from the h = 2 case
// (profiles/README.md, "Wider states"):
//   * the gather ran as rowptr -> (col, val) -> feature-row chains, three dependent memory rounds per operator
//     and row, with 8 warps per SM (one CTA: the weight block and the tile fill shared memory): latency-bound;
//   * the linear ran as fp32 FMAs fed from shared memory (8 LDS.128 per 64 FMA: shared-memory bound).
// Here
//   * a CTA has 16 warps and STAGES the CSR structure of its 64-row tile first: the row pointers of every CSR
//     operator (one coalesced round), then the tile's contiguous slice of (col, val) (a second coalesced round)
//     into shared memory.  The gather proper then issues nothing but independent 16-byte feature loads: one item
//     per thread and tile = two chunks of a row, own row and 3-4 neighbour rows in flight together, raw rows
//     summed and the producer's batch-norm applied once per sum;
//   * the three contractions (forward Z = T W, backward gX = T W^T-block, dW += T^T Xn) run on the tensor
//     cores as mma.sync m16n8k8 TF32 with the 3xTF32 error-compensated split (a = hi + lo: lo*hi + hi*lo + hi*hi).
//     The tensor core accumulates with truncation, so the large terms are summed outside it in round-to-nearest
//     fp32 (see mma_3xtf32_pair); that keeps the 1e-4 parity bound of fp32 with margin (tests/test_wide_dispatch.py
//     emulates the arithmetic).  dW lives in accumulator fragments across all tiles of the CTA and is flushed once;
//   * in the backward every CTA works through its share of the self tiles and then of the cross tiles.
// Long rows (> ENG_LONG_ROW entries) and the run-length ranges of the transposed line-graph operator keep using
// gather_deferred of engine.cu.  Semantics are those of fwd_kernel / bwd_kernel (layers_mnb.py:189-225,
// batch_normalization.py:34-43,65-77).  Measurements and the ncu evidence behind each choice: profiles/README.md.
#pragma once

#define WD_THREADS 512
#define WD_WARPS 16
#define WD_TR 64
#define WD_CAP 1024        // staged CSR entries per operator and tile
#define WD_SLOTS 2         // staged CSR operators of the self part (forward: plus the Pm/Pd pattern)
#define WD_MAXDW 4         // 16x16 dW blocks per warp

#include "engine_mma.cuh"      // tf32_split, mma_tf32, the 3xTF32 steps

// One warp: acc[j] (j = 0, 1: two 16x8 blocks side by side) = A[16 x Kd] * B[Kd x 16],
// A(m, k) = A[m * lda + k] (lda = 4 mod 8: conflict-free), B(k, n) = B[k * ldb + n] (ldb = 8 mod 16).
__device__ __forceinline__ void wide_mma_rows(float (&acc)[2][4], const float* __restrict__ A, int lda,
                                              const float* __restrict__ B, int ldb, int Kd) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* A0 = A + g * lda + t;         // rows g / g + 8 of the block
    const float* A1 = A0 + 8 * lda;
    const float* B0 = B + t * ldb + g;         // k rows t / t + 4 of the k-step
    const float* B1 = B0 + 4 * ldb;
    const int bstep = 8 * ldb;
    float sm[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = sm[j][e] = 0.f;
    int k0 = 0;
    for (; k0 + 16 <= Kd; k0 += 16) {
        uint32_t ah[4], al[4], ah2[4], al2[4];
        tf32_split(A0[k0], ah[0], al[0]);
        tf32_split(A1[k0], ah[1], al[1]);
        tf32_split(A0[k0 + 4], ah[2], al[2]);
        tf32_split(A1[k0 + 4], ah[3], al[3]);
        tf32_split(A0[k0 + 8], ah2[0], al2[0]);
        tf32_split(A1[k0 + 8], ah2[1], al2[1]);
        tf32_split(A0[k0 + 12], ah2[2], al2[2]);
        tf32_split(A1[k0 + 12], ah2[3], al2[3]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            mma_3xtf32_pair(acc[j], sm[j], ah, al, split_b(B0[j * 8], B1[j * 8]), ah2, al2,
                            split_b(B0[bstep + j * 8], B1[bstep + j * 8]));
        B0 += 2 * bstep;
        B1 += 2 * bstep;
    }
    if (k0 < Kd) {                             // Kd = 8 mod 16
        uint32_t ah[4], al[4];
        tf32_split(A0[k0], ah[0], al[0]);
        tf32_split(A1[k0], ah[1], al[1]);
        tf32_split(A0[k0 + 4], ah[2], al[2]);
        tf32_split(A1[k0 + 4], ah[3], al[3]);
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_3xtf32(acc[j], sm[j], ah, al, split_b(B0[j * 8], B1[j * 8]));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] += sm[j][e];
}

// column sums of a 16 x 8 accumulator block: the lanes that share t = lane % 4 hold the same two columns; their
// partial sums (s: plain, q: second statistic) go to the CTA's fp64 totals stat[col], stat[F + col]
__device__ __forceinline__ void wide_stat_flush(double* stat, int F, int col, float s0, float s1, float q0, float q1) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    if ((threadIdx.x & 31) < 4) {
        atomicAdd(stat + col, (double)s0);
        atomicAdd(stat + col + 1, (double)s1);
        atomicAdd(stat + F + col, (double)q0);
        atomicAdd(stat + F + col + 1, (double)q1);
    }
}

// ---- staged CSR structure of one tile ---------------------------------------------------------------
struct WideStage {
    int rp[WD_SLOTS + 1][WD_TR + 1];   // row pointers of the tile's rows (+1)
    int base[WD_SLOTS + 1];            // rp[s][0]
    int staged[WD_SLOTS + 1];          // 1 = (col, val) of the tile are in shared memory
    int slot_of[HGNN_MAX_OPS];         // stage slot of operator t, or -1
    int op_of[WD_SLOTS];               // operator of slot s, or -1
};

// Entries of one CSR row, either from the staged slice or from global memory.
struct WideRow {
    const int* col;      // indexable by k in [k0, k1)
    const float* val;
    const float* val2;   // second value array (Pm / Pd share one pattern), or NULL
    int k0, k1;
    bool smem;
};
__device__ __forceinline__ int wr_col(const WideRow& w, int k) { return w.smem ? w.col[k] : __ldg(w.col + k); }
__device__ __forceinline__ float wr_val(const WideRow& w, int k) { return w.smem ? w.val[k] : __ldg(w.val + k); }
__device__ __forceinline__ float wr_val2(const WideRow& w, int k) { return w.smem ? w.val2[k] : __ldg(w.val2 + k); }

// acc[h] = sum_k val[k] * ld(col[k], xo + h * xs), h < NCH: BATCH * NCH independent 16-byte feature loads in flight
// (raw rows; the loader's fix-up of the weighted sum is applied once at the end).  An item of the tile gathers covers
// NCH = 2 chunks of its row, half a row apart, so the index work (row pointers, entries, predicates) is paid once
// per 32 bytes of every gathered row and a tile is one item per thread.
template <int BATCH, int NCH, typename L>
__device__ __forceinline__ void wide_gather(const WideRow& w, const L& ld_, int xo, int xs, V<4> (&acc)[NCH]) {
#pragma unroll
    for (int h = 0; h < NCH; ++h) acc[h] = V<4>::zero();
    float vsum = 0.f;
    for (int k = w.k0; k < w.k1; k += BATCH) {
        int c[BATCH];
        float v[BATCH];
        V<4> x[BATCH][NCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const bool on = k + j < w.k1;
            const int kk = on ? k + j : k;
            c[j] = wr_col(w, kk);
            v[j] = on ? wr_val(w, kk) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
#pragma unroll
            for (int h = 0; h < NCH; ++h) {
                x[j][h] = V<4>::zero();
                if (k + j < w.k1) x[j][h] = ld_.raw(c[j], xo + h * xs);
            }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
#pragma unroll
            for (int h = 0; h < NCH; ++h) acc[h].fma(v[j], x[j][h]);
            vsum += v[j];
        }
    }
#pragma unroll
    for (int h = 0; h < NCH; ++h) ld_.finish(acc[h], vsum, xo + h * xs);
}

// Stage the structure of rows [row0, row0 + trc) of the CSR operators in `slots` (all threads; two barriers inside,
// ends synced).  slot s < WD_SLOTS: operator st->op_of[s] of `ops`; slot WD_SLOTS: the (rowptr, col, v1, v2) pattern
// of the cross part when p_rowptr != NULL.  Also pushes the rows that own run-length entries to dl->rng_items.
template <int CAP = WD_CAP>
__device__ __forceinline__ void wide_stage(WideStage* st, const OpList& ops, const int* __restrict__ p_rowptr,
                                           const int* __restrict__ p_col, const float* __restrict__ p_v1,
                                           const float* __restrict__ p_v2, int row0, int trc, int* scol, float* sval,
                                           int* pcol, float* pv1, float* pv2, DeferList* dl) {
    const int tid = threadIdx.x;
    for (int i = tid; i < (WD_SLOTS + 1) * (WD_TR + 1); i += WD_THREADS) {
        const int s = i / (WD_TR + 1), r = i - s * (WD_TR + 1);
        const int* rp = nullptr;
        if (s < WD_SLOTS) { if (st->op_of[s] >= 0) rp = ops.rowptr[st->op_of[s]]; }
        else rp = p_rowptr;
        if (rp && r <= trc) st->rp[s][r] = __ldg(rp + row0 + r);
    }
    // rows with a run-length part: one push per (row, operator), as gather_or_defer does
    for (int i = tid; i < ops.n * WD_TR; i += WD_THREADS) {
        const int t = i / WD_TR, r = i - t * WD_TR;
        if (r < trc && ops.kind[t] == HGNN_OP_CSR && ops.rng_rowptr[t] &&
            __ldg(ops.rng_rowptr[t] + row0 + r + 1) > __ldg(ops.rng_rowptr[t] + row0 + r)) {
            const int slot = atomicAdd(&dl->rng_cnt, 1);
            if (slot < ENG_THREADS) dl->rng_items[slot] = defer_code(t, 0, r);
            else __trap();
        }
    }
    __syncthreads();
    for (int s = 0; s <= WD_SLOTS; ++s) {
        const bool cross = s == WD_SLOTS;
        const bool have = cross ? p_rowptr != nullptr : st->op_of[s] >= 0;
        if (!have) continue;
        const int b = st->rp[s][0], n = st->rp[s][trc] - b;
        const bool fits = n <= CAP;
        if (tid == 0) { st->base[s] = b; st->staged[s] = fits ? 1 : 0; }
        if (!fits) continue;
        if (cross) {
            for (int i = tid; i < n; i += WD_THREADS) {
                pcol[i] = __ldg(p_col + b + i);
                pv1[i] = __ldg(p_v1 + b + i);
                pv2[i] = __ldg(p_v2 + b + i);
            }
        } else {
            const int t = st->op_of[s];
            const int* __restrict__ gc = ops.col[t];
            const float* __restrict__ gv = ops.val[t];
            for (int i = tid; i < n; i += WD_THREADS) {
                scol[s * CAP + i] = __ldg(gc + b + i);
                sval[s * CAP + i] = __ldg(gv + b + i);
            }
        }
    }
    __syncthreads();
}

// (k0, k1, arrays) of row r of the tile for operator t
template <int CAP = WD_CAP>
__device__ __forceinline__ WideRow wide_row(const WideStage* st, const OpList& ops, int t, int row0, int r,
                                            const int* scol, const float* sval) {
    WideRow w;
    w.val2 = nullptr;
    const int s = st->slot_of[t];
    if (s >= 0) {
        w.k0 = st->rp[s][r];
        w.k1 = st->rp[s][r + 1];
        if (st->staged[s]) {
            w.smem = true;
            w.col = scol + s * CAP - st->base[s];
            w.val = sval + s * CAP - st->base[s];
            return w;
        }
    } else {
        w.k0 = __ldg(ops.rowptr[t] + row0 + r);
        w.k1 = __ldg(ops.rowptr[t] + row0 + r + 1);
    }
    w.smem = false;
    w.col = ops.col[t];
    w.val = ops.val[t];
    return w;
}

__device__ __forceinline__ void wide_assign_slots(WideStage* st, const OpList& ops) {
    if (threadIdx.x == 0) {
        int ns = 0;
        for (int s = 0; s < WD_SLOTS; ++s) st->op_of[s] = -1;
        for (int t = 0; t < HGNN_MAX_OPS; ++t) {
            st->slot_of[t] = -1;
            if (t < ops.n && ops.kind[t] == HGNN_OP_CSR && ns < WD_SLOTS) {
                st->slot_of[t] = ns;
                st->op_of[ns] = t;
                ++ns;
            }
        }
    }
}

// Two value arrays on one pattern (Pm / Pd): am = sum_k val[k] x_k, ad = sum_k val2[k] x_k, every row loaded once
template <int BATCH, int NCH, typename L>
__device__ __forceinline__ void wide_gather2(const WideRow& w, const L& ld_, int xo, int xs, V<4> (&am)[NCH],
                                             V<4> (&ad)[NCH]) {
#pragma unroll
    for (int h = 0; h < NCH; ++h) { am[h] = V<4>::zero(); ad[h] = V<4>::zero(); }
    float sm_ = 0.f, sd_ = 0.f;
    for (int k = w.k0; k < w.k1; k += BATCH) {
        int c[BATCH];
        float vm[BATCH], vd[BATCH];
        V<4> x[BATCH][NCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const bool on = k + j < w.k1;
            const int kk = on ? k + j : k;
            c[j] = wr_col(w, kk);
            vm[j] = on ? wr_val(w, kk) : 0.f;
            vd[j] = on ? wr_val2(w, kk) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
#pragma unroll
            for (int h = 0; h < NCH; ++h) {
                x[j][h] = V<4>::zero();
                if (k + j < w.k1) x[j][h] = ld_.raw(c[j], xo + h * xs);
            }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
#pragma unroll
            for (int h = 0; h < NCH; ++h) {
                am[h].fma(vm[j], x[j][h]);
                ad[h].fma(vd[j], x[j][h]);
            }
            sm_ += vm[j];
            sd_ += vd[j];
        }
    }
#pragma unroll
    for (int h = 0; h < NCH; ++h) {
        ld_.finish(am[h], sm_, xo + h * xs);
        ld_.finish(ad[h], sd_, xo + h * xs);
    }
}

// Self part of the tile gather: item (r, q) = feature chunks q and q + F/8 of row r (F is a multiple of 8); writes
// tile[r][t * F + ..] for every operator t.  CSR operators first (their feature loads are issued together with the
// own-row load), identity / diagonal blocks last.  `dual`: the list is two CSR operators on ONE pattern (Pm^T / Pd^T
// of the backward's cross part): one pass over the entries, every gathered row loaded once.
template <int BATCH, typename L>
__device__ __forceinline__ void wide_gather_self(const WideStage* st, const OpList& ops, const L& ld_, int F, int row0,
                                                 int trc, float* tile, int ldt, const int* scol, const float* sval,
                                                 DeferList* dl, bool own_block, int own_col, bool dual) {
    const int Qh = F >> 3, xs = F >> 1, K = ops.n;
    bool need_own = own_block;
    for (int t = 0; t < K; ++t) need_own = need_own || ops.kind[t] != HGNN_OP_CSR;
    for (int i = threadIdx.x; i < Qh * trc; i += WD_THREADS) {
        const int r = i / Qh, q = i - r * Qh;
        const int row = row0 + r, xo = q << 2;
        float* trow = tile + r * ldt;
        if (dual) {
            WideRow w = wide_row(st, ops, 0, row0, r, scol, sval);
            const WideRow w1 = wide_row(st, ops, 1, row0, r, scol, sval);
            w.val2 = w1.val;
            V<4> am[2], ad[2];
            wide_gather2<BATCH, 2>(w, ld_, xo, xs, am, ad);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                am[h].store(trow + xo + h * xs);
                ad[h].store(trow + F + xo + h * xs);
            }
            continue;
        }
        V<4> own[2];
        own[0] = own[1] = V<4>::zero();
        if (need_own) {
            own[0] = ld_(row, xo);
            own[1] = ld_(row, xo + xs);
        }
        for (int t = 0; t < K; ++t) {
            if (ops.kind[t] != HGNN_OP_CSR) continue;
            const WideRow w = wide_row(st, ops, t, row0, r, scol, sval);
            if (w.k1 - w.k0 > ENG_LONG_ROW) {
                const int slot = atomicAdd(&dl->cnt, 2);
                if (slot + 1 < ENG_MAX_DEFER) {
                    dl->items[slot] = defer_code(t, q, r);
                    dl->items[slot + 1] = defer_code(t, q + Qh, r);
                    continue;
                }
            }
            V<4> acc[2];
            wide_gather<BATCH, 2>(w, ld_, xo, xs, acc);
            acc[0].store(trow + t * F + xo);
            acc[1].store(trow + t * F + xo + xs);
        }
        for (int t = 0; t < K; ++t) {
            if (ops.kind[t] == HGNN_OP_IDENT) {
                own[0].store(trow + t * F + xo);
                own[1].store(trow + t * F + xo + xs);
            } else if (ops.kind[t] == HGNN_OP_DIAG) {
                const float dg = __ldg(ops.diag[t] + row);
                V<4> x0 = own[0], x1 = own[1];
                x0.scale(dg);
                x1.scale(dg);
                x0.store(trow + t * F + xo);
                x1.store(trow + t * F + xo + xs);
            }
        }
        if (own_block) {
            own[0].store(trow + own_col + xo);
            own[1].store(trow + own_col + xo + xs);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct WideFwdLayout { int ldw, Wt, bias, sc_s, sh_s, sc_c, sh_c, tile, stage, scol, sval, pcol, pv1, pv2, total; };
__host__ __device__ inline WideFwdLayout wide_fwd_layout(int Cin, int Cp, int Fout, int Fs, int Fc, int TR) {
    WideFwdLayout l;
    int o = 0;
    l.ldw = Fout + 8;
    l.Wt = o; o += Cin * l.ldw;
    l.bias = o; o += Fout;
    l.sc_s = o; o += Fs;  l.sh_s = o; o += Fs;
    l.sc_c = o; o += Fc;  l.sh_c = o; o += Fc;
    o = (o + 3) & ~3;
    l.tile = o; o += TR * Cp;
    l.stage = o; o += (int)((sizeof(WideStage) + 15) / 16) * 4;
    l.scol = o; o += WD_SLOTS * WD_CAP;
    l.sval = o; o += WD_SLOTS * WD_CAP;
    l.pcol = o; o += WD_CAP;
    l.pv1 = o; o += WD_CAP;
    l.pv2 = o; o += WD_CAP;
    l.total = o;
    return l;
}

__global__ void __launch_bounds__(WD_THREADS, 1)
fwd_wide_kernel(const FwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double dscratch[WD_THREADS];
    __shared__ double dtot[256];
    __shared__ DeferList dl;
    __shared__ float wpart[4 * WD_WARPS];
    const int Cin = a.Cin, Cp = a.Cin_pad, Fout = a.Fout, TR = a.TR;
    const int K = a.ops.n, Fs = a.Fs, Fc = a.Fc;
    const WideFwdLayout lay = wide_fwd_layout(Cin, Cp, Fout, Fs, Fc, TR);
    const int ldw = lay.ldw;
    float* Wt = smem + lay.Wt;                 // [Cin][ldw]
    float* bias = smem + lay.bias;
    float* sc_s = smem + lay.sc_s;
    float* sh_s = smem + lay.sh_s;
    float* sc_c = smem + lay.sc_c;
    float* sh_c = smem + lay.sh_c;
    float* tile = smem + lay.tile;             // [TR][Cp]
    WideStage* st = reinterpret_cast<WideStage*>(smem + lay.stage);
    int* scol = reinterpret_cast<int*>(smem + lay.scol);
    float* sval = smem + lay.sval;
    int* pcol = reinterpret_cast<int*>(smem + lay.pcol);
    float* pv1 = smem + lay.pv1;
    float* pv2 = smem + lay.pv2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;

    for (int i = tid; i < Cin * Fout; i += WD_THREADS) {
        const int o = i / Cin, c = i - o * Cin;
        Wt[c * ldw + o] = (o < a.Ha) ? a.Wa[(size_t)o * Cin + c] : a.Wb[(size_t)(o - a.Ha) * Cin + c];
    }
    for (int o = tid; o < Fout; o += WD_THREADS)
        bias[o] = (o < a.Ha) ? (a.ba ? a.ba[o] : 0.f) : (a.bb ? a.bb[o - a.Ha] : 0.f);
    if (tid == 0) dl.rsum_id = -1;
    wide_assign_slots(st, a.ops);
    const bool cross = a.p_rowptr != nullptr;
    const bool aff_s = bn_vectors(a.bn_s, Fs, sc_s, sh_s, nullptr, nullptr, dtot, dscratch);
    const bool aff_c = cross ? bn_vectors(a.bn_c, Fc, sc_c, sh_c, nullptr, nullptr, dtot, dscratch) : false;
    const AffineLoader<4> ls{a.Xs, Fs, sc_s, sh_s, aff_s};
    const AffineLoader<4> lc{a.Xc, Fc, sc_c, sh_c, aff_c};
    __syncthreads();
    double* sstat = dtot;                      // (sum z, sum z^2) of this CTA's rows, [2 * Fout]
    for (int i = tid; i < 2 * Fout; i += WD_THREADS) sstat[i] = 0.0;

    const int xc0 = K * Fs;
    const int Qc = Fc >> 3;
    const int MB = TR >> 4, NJ = Fout >> 4;
    const int ntiles = (a.R + TR - 1) / TR;
    const bool one_block = MB * NJ == WD_WARPS;
    float rst[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) rst[j][0] = rst[j][1] = rst[j][2] = rst[j][3] = 0.f;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, a.R - row0);
        if (tid == 0) { dl.cnt = 0; dl.rng_cnt = 0; }
        __syncthreads();                       // previous tile's contraction is done with `tile`
        wide_stage(st, a.ops, a.p_rowptr, a.p_col, a.p_pm, a.p_pd, row0, trc, scol, sval, pcol, pv1, pv2, &dl);
        wide_gather_self<4>(st, a.ops, ls, Fs, row0, trc, tile, Cp, scol, sval, &dl, false, 0, false);
        if (cross) {
            const bool sm = st->staged[WD_SLOTS] != 0;
            const int base = st->base[WD_SLOTS];
            for (int i = tid; i < Qc * trc; i += WD_THREADS) {      // Qc = Fc / 8: two chunks per item
                const int r = i / Qc, q = i - r * Qc;
                const int xo = q << 2, xs = Fc >> 1;
                WideRow w;
                w.k0 = st->rp[WD_SLOTS][r];
                w.k1 = st->rp[WD_SLOTS][r + 1];
                w.smem = sm;
                w.col = sm ? pcol - base : a.p_col;
                w.val = sm ? pv1 - base : a.p_pm;
                w.val2 = sm ? pv2 - base : a.p_pd;
                V<4> am[2], ad[2];
                wide_gather2<4, 2>(w, lc, xo, xs, am, ad);
                float* trow = tile + r * Cp;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    am[h].store(trow + xc0 + xo + h * xs);
                    ad[h].store(trow + xc0 + Fc + xo + h * xs);
                }
            }
        }
        __syncthreads();
        gather_deferred<4>(a.ops, &dl, row0, ls, Fs, tile, Cp, wpart);
        __syncthreads();
        // ---- Z = relu(T W + b), statistics of Z: one 16 x 16 block per warp and pass
        for (int b = warp; b < MB * NJ; b += WD_WARPS) {
            const int mi = b % MB, nj = b / MB;
            float acc[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            wide_mma_rows(acc, tile + (mi << 4) * Cp, Cp, Wt + (nj << 4), ldw, Cin);
            const int rA = (mi << 4) + g, rB = rA + 8;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = (nj << 4) + (j << 3) + (t4 << 1);
                const float b0 = bias[col], b1 = bias[col + 1];
                float v0 = acc[j][0] + b0, v1 = acc[j][1] + b1, v2 = acc[j][2] + b0, v3 = acc[j][3] + b1;
                if (col >= a.relu_from) { v0 = fmaxf(v0, 0.f); v2 = fmaxf(v2, 0.f); }
                if (col + 1 >= a.relu_from) { v1 = fmaxf(v1, 0.f); v3 = fmaxf(v3, 0.f); }
                if (rA < trc) *reinterpret_cast<float2*>(a.Z + (size_t)(row0 + rA) * Fout + col) = make_float2(v0, v1);
                else v0 = v1 = 0.f;
                if (rB < trc) *reinterpret_cast<float2*>(a.Z + (size_t)(row0 + rB) * Fout + col) = make_float2(v2, v3);
                else v2 = v3 = 0.f;
                if (a.acc_out) {
                    const float s0 = v0 + v2, s1 = v1 + v3;
                    const float q0 = fmaf(v0, v0, v2 * v2), q1 = fmaf(v1, v1, v3 * v3);
                    if (one_block) {           // this warp's columns never change: keep the partial sums in registers
                        rst[j][0] += s0; rst[j][1] += s1; rst[j][2] += q0; rst[j][3] += q1;
                    } else {
                        wide_stat_flush(sstat, Fout, col, s0, s1, q0, q1);
                    }
                }
            }
        }
    }
    if (a.acc_out && one_block) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
            wide_stat_flush(sstat, Fout, ((warp / MB) << 4) + (j << 3) + (t4 << 1), rst[j][0], rst[j][1], rst[j][2], rst[j][3]);
    }
    if (a.acc_out) {
        __syncthreads();
        const int nb = hgnn_ws_bins(2 * Fout);
        for (int i = tid; i < 2 * Fout; i += WD_THREADS) accum_add(a.acc_out, 2 * Fout, nb, i, sstat[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct WideBwdLayout { int ldw, Wsm, sc, sh, mu, rs, tile, xt, stage, scol, sval, total; };
__host__ __device__ inline WideBwdLayout wide_bwd_layout(int nT, int Fx, int Tp, int Xp, int TR) {
    WideBwdLayout l;
    int o = 0;
    l.ldw = Fx + 8;
    l.Wsm = o; o += nT * l.ldw;
    l.sc = o; o += Fx;  l.sh = o; o += Fx;  l.mu = o; o += Fx;  l.rs = o; o += Fx;
    o = (o + 3) & ~3;
    l.tile = o; o += TR * Tp;
    l.xt = o; o += TR * Xp;
    l.stage = o; o += (int)((sizeof(WideStage) + 15) / 16) * 4;
    l.scol = o; o += WD_SLOTS * WD_CAP;
    l.sval = o; o += WD_SLOTS * WD_CAP;
    l.total = o;
    return l;
}

__device__ __forceinline__ void bwd_wide_part(const BwdArgs& a, const BwdPart& p, bool is_self, int first_tile,
                                              int tile_stride, float* smem, double* dscratch, double* dtot,
                                              DeferList* dl, float* wpart, const float* c0, const float* c1,
                                              const float* c2, bool has_bn) {
    const int nT = p.nT, Tp = p.Tp, Fx = p.Fx, Xp = p.Xp, Fg = a.Fg, TR = p.TR;
    const WideBwdLayout lay = wide_bwd_layout(nT, Fx, Tp, Xp, TR);
    const int ldw = lay.ldw;
    float* Wsm = smem + lay.Wsm;               // [nT][ldw]: Wsm[c][f] = W[o(c)][col0 + t(c) * Fx + f]
    float* sc = smem + lay.sc;
    float* sh = smem + lay.sh;
    float* mu = smem + lay.mu;
    float* rs = smem + lay.rs;
    float* tile = smem + lay.tile;             // [TR][Tp]: gathered gPre blocks (+ own gPre row for the self part)
    float* xt = smem + lay.xt;                 // [TR][Xp]: raw rows of the input
    WideStage* st = reinterpret_cast<WideStage*>(smem + lay.stage);
    int* scol = reinterpret_cast<int*>(smem + lay.scol);
    float* sval = smem + lay.sval;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const bool want_dw = a.dW_bins != nullptr;

    __syncthreads();                           // a previous part of this CTA is done with shared memory
    if (tid == 0) dl->rsum_id = -1;
    for (int i = tid; i < nT * Fx; i += WD_THREADS) {
        const int c = i / Fx, f = i - c * Fx;
        const int t = c / Fg, o = c - t * Fg;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        Wsm[c * ldw + f] = wrow[p.col0 + t * Fx + f];
    }
    wide_assign_slots(st, p.ops);
    const bool x_aff = bn_vectors(p.bn, Fx, sc, sh, mu, rs, dtot, dscratch);
    (void)x_aff;                               // sc = 1, sh = 0 when the input was not normalised
    const GpreLoader<4> lg{a.gY, a.Z, Fg, c0, c1, c2, a.relu_from, has_bn};
    const bool dual = !is_self && p.ops.n == 2 && p.ops.kind[0] == HGNN_OP_CSR && p.ops.kind[1] == HGNN_OP_CSR &&
                      p.ops.rowptr[0] == p.ops.rowptr[1] && p.ops.col[0] == p.ops.col[1] &&
                      !p.ops.rng_rowptr[0] && !p.ops.rng_rowptr[1];
    __syncthreads();
    double* sstat = dtot;                      // (sum g, sum g * xhat) of the rows produced here, [2 * Fx]
    double* sdb = dtot + 2 * Fx;               // dbias partial sums [Fg]   (2 Fx + Fg <= 384 <= 512)
    for (int i = tid; i < 2 * Fx + Fg; i += WD_THREADS) dtot[i] = 0.0;

    const int MBr = TR >> 4, NJx = Fx >> 4;    // gX blocks
    // dW = MBc x NJx blocks of 16 x 16: warp w owns column block nj = w % NJx and the row blocks
    // mi = w / NJx + i * (WD_WARPS / NJx), i < WD_MAXDW, so that its blocks share one B fragment per k-step
    const int MBc = nT >> 4;
    const int dw_nj = warp % NJx, dw_m0 = warp / NJx, dw_ms = WD_WARPS / NJx;
    float dw[WD_MAXDW][2][4];
#pragma unroll
    for (int i = 0; i < WD_MAXDW; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dw[i][j][0] = dw[i][j][1] = dw[i][j][2] = dw[i][j][3] = 0.f;
    const bool one_block = MBr * NJx == WD_WARPS;
    float rst[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) rst[j][0] = rst[j][1] = rst[j][2] = rst[j][3] = 0.f;
    const int db_groups = WD_THREADS / Fg;
    const int db_o = tid % Fg, db_g = tid / Fg;
    float dbacc = 0.f;

    for (int tile_id = first_tile; tile_id < p.tiles; tile_id += tile_stride) {
        const int row0 = tile_id * TR;
        const int trc = min(TR, p.R - row0);
        if (tid == 0) { dl->cnt = 0; dl->rng_cnt = 0; }
        __syncthreads();                       // previous tile's contractions are done with tile / xt
        wide_stage(st, p.ops, nullptr, nullptr, nullptr, nullptr, row0, trc, scol, sval, nullptr, nullptr, nullptr, dl);
        wide_gather_self<3>(st, p.ops, lg, Fg, row0, trc, tile, Tp, scol, sval, dl, is_self, nT, dual);
        {
            const int NQ = Fx >> 2;
            for (int i = tid; i < trc * NQ; i += WD_THREADS) {
                const int r = i / NQ, g4 = i - r * NQ;
                *reinterpret_cast<float4*>(xt + r * Xp + g4 * 4) =
                    __ldg(reinterpret_cast<const float4*>(p.X + (size_t)row0 * Fx) + i);
            }
        }
        if (trc < TR) {                        // the dW contraction runs over all TR rows: clear the unused ones
            for (int i = tid; i < (TR - trc) * Tp; i += WD_THREADS) tile[trc * Tp + i] = 0.f;
            for (int i = tid; i < (TR - trc) * Xp; i += WD_THREADS) xt[trc * Xp + i] = 0.f;
        }
        __syncthreads();
        gather_deferred<4>(p.ops, dl, row0, lg, Fg, tile, Tp, wpart);
        __syncthreads();
        // ---- gX = T Wsm  (+ statistics of what was produced, for the input's own BN backward)
        if (p.gX) {
            for (int b = warp; b < MBr * NJx; b += WD_WARPS) {
                const int mi = b % MBr, nj = b / MBr;
                float acc[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
                wide_mma_rows(acc, tile + (mi << 4) * Tp, Tp, Wsm + (nj << 4), ldw, nT);
                const int rA = (mi << 4) + g, rB = rA + 8;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int f = (nj << 4) + (j << 3) + (t4 << 1);
                    float v0 = acc[j][0], v1 = acc[j][1], v2 = acc[j][2], v3 = acc[j][3];
                    if (rA >= trc) v0 = v1 = 0.f;
                    if (rB >= trc) v2 = v3 = 0.f;
                    if (p.acc_b) {
                        const float m0 = mu[f], m1 = mu[f + 1], r0 = rs[f], r1 = rs[f + 1];
                        const float2 xa = *reinterpret_cast<const float2*>(xt + rA * Xp + f);
                        const float2 xb = *reinterpret_cast<const float2*>(xt + rB * Xp + f);
                        const float s0 = v0 + v2, s1 = v1 + v3;
                        const float q0 = fmaf(v0, (xa.x - m0) * r0, v2 * ((xb.x - m0) * r0));
                        const float q1 = fmaf(v1, (xa.y - m1) * r1, v3 * ((xb.y - m1) * r1));
                        if (one_block) {
                            rst[j][0] += s0; rst[j][1] += s1; rst[j][2] += q0; rst[j][3] += q1;
                        } else {
                            wide_stat_flush(sstat, Fx, f, s0, s1, q0, q1);
                        }
                    }
                    if (rA < trc) {
                        float2* dst = reinterpret_cast<float2*>(p.gX + (size_t)(row0 + rA) * Fx + f);
                        float2 o2 = make_float2(v0, v1);
                        if (p.accumulate) { const float2 old = *dst; o2.x += old.x; o2.y += old.y; }
                        *dst = o2;
                    }
                    if (rB < trc) {
                        float2* dst = reinterpret_cast<float2*>(p.gX + (size_t)(row0 + rB) * Fx + f);
                        float2 o2 = make_float2(v2, v3);
                        if (p.accumulate) { const float2 old = *dst; o2.x += old.x; o2.y += old.y; }
                        *dst = o2;
                    }
                }
            }
        }
        // ---- dW[c][f] += sum_r T[r][c] * xnorm[r][f]  (A = T^T from the tile, B = normalised input rows)
        if (want_dw) {
            const float* Bp = xt + t4 * Xp + (dw_nj << 4) + g;
            const float s0 = sc[(dw_nj << 4) + g], h0 = sh[(dw_nj << 4) + g];
            const float s1 = sc[(dw_nj << 4) + 8 + g], h1 = sh[(dw_nj << 4) + 8 + g];
            const float* Ap = tile + t4 * Tp + g;
            for (int k0 = 0; k0 < TR; k0 += 16) {
                const float* b = Bp + k0 * Xp;
                const SplitB b00 = split_b(fmaf(b[0], s0, h0), fmaf(b[4 * Xp], s0, h0));
                const SplitB b01 = split_b(fmaf(b[8], s1, h1), fmaf(b[4 * Xp + 8], s1, h1));
                const SplitB b10 = split_b(fmaf(b[8 * Xp], s0, h0), fmaf(b[12 * Xp], s0, h0));
                const SplitB b11 = split_b(fmaf(b[8 * Xp + 8], s1, h1), fmaf(b[12 * Xp + 8], s1, h1));
#pragma unroll
                for (int i = 0; i < WD_MAXDW; ++i) {
                    const int mi = dw_m0 + i * dw_ms;
                    if (mi < MBc) {
                        const float* a = Ap + k0 * Tp + (mi << 4);
                        uint32_t ah[4], al[4], ah2[4], al2[4];
                        tf32_split(a[0], ah[0], al[0]);
                        tf32_split(a[8], ah[1], al[1]);
                        tf32_split(a[4 * Tp], ah[2], al[2]);
                        tf32_split(a[4 * Tp + 8], ah[3], al[3]);
                        tf32_split(a[8 * Tp], ah2[0], al2[0]);
                        tf32_split(a[8 * Tp + 8], ah2[1], al2[1]);
                        tf32_split(a[12 * Tp], ah2[2], al2[2]);
                        tf32_split(a[12 * Tp + 8], ah2[3], al2[3]);
                        mma_3xtf32_fresh_pair(dw[i][0], ah, al, b00, ah2, al2, b10);
                        mma_3xtf32_fresh_pair(dw[i][1], ah, al, b01, ah2, al2, b11);
                    }
                }
            }
            if (is_self && db_g < db_groups) {
                for (int r = db_g; r < trc; r += db_groups) dbacc += tile[r * Tp + nT + db_o];
            }
        }
    }
    // ---- flush the CTA's partial sums to the binned fp64 accumulators
    if (want_dw) {
        const int nbw = hgnn_ws_bins(Fg * a.Cin);
#pragma unroll
        for (int i = 0; i < WD_MAXDW; ++i) {
            const int mi = dw_m0 + i * dw_ms, nj = dw_nj;
            if (mi < MBc) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = (mi << 4) + g + ((e & 2) ? 8 : 0);
                        const int f = (nj << 4) + (j << 3) + (t4 << 1) + (e & 1);
                        const int t = c / Fg, o = c - t * Fg;
                        accum_add(a.dW_bins, Fg * a.Cin, nbw, o * a.Cin + p.col0 + t * Fx + f, (double)dw[i][j][e]);
                    }
            }
        }
        if (is_self && a.db_bins) {
            if (db_g < db_groups) atomicAdd(sdb + db_o, (double)dbacc);
            __syncthreads();
            const int nbb = hgnn_ws_bins(Fg);
            for (int o = tid; o < Fg; o += WD_THREADS) accum_add(a.db_bins, Fg, nbb, o, sdb[o]);
        }
    }
    if (p.acc_b && p.gX) {
        if (one_block) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                wide_stat_flush(sstat, Fx, ((warp / MBr) << 4) + (j << 3) + (t4 << 1), rst[j][0], rst[j][1], rst[j][2], rst[j][3]);
        }
        __syncthreads();
        const int nb = hgnn_ws_bins(2 * Fx);
        for (int i = tid; i < 2 * Fx; i += WD_THREADS) accum_add(p.acc_b, 2 * Fx, nb, i, sstat[i]);
    }
}

__global__ void __launch_bounds__(WD_THREADS, 1)
bwd_wide_kernel(const BwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ double dscratch[WD_THREADS];
    __shared__ double dtot[512];
    __shared__ DeferList dl;
    __shared__ float wpart[4 * WD_WARPS];
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
        for (int f = threadIdx.x; f < Fg; f += WD_THREADS) {
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
    // Every CTA takes its share of BOTH parts, one after the other (self tiles dealt from CTA 0 up, cross tiles
    // from the last CTA down, so the remainders land on different CTAs): the two kinds of tile cost different
    // amounts, and a split of the CTAs by tile count left the cross CTAs of the edge side running 25 % longer than
    // the self CTAs (ncu sampling, profiles/README.md).  Costs one more prologue and dW flush per CTA.
    if (a.self.R > 0)
        bwd_wide_part(a, a.self, true, blockIdx.x, gridDim.x, smem, dscratch, dtot, &dl, wpart, c0, c1, c2, has_bn);
    if (a.cross.R > 0)
        bwd_wide_part(a, a.cross, false, gridDim.x - 1 - blockIdx.x, gridDim.x, smem, dscratch, dtot, &dl, wpart, c0, c1,
                      c2, has_bn);
}
