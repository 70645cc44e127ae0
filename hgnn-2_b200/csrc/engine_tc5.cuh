// engine_tc5.cuh -- EXPERIMENTAL (off unless HGNN_B200_WIDE_TC5=1; written at the end of round 1 without GPU time
// left: it compiles to UTCHMMA / LDTM SASS but has NOT run yet - bring-up order in DESIGN.md 3b, stand-alone probe
// profiles/tc5_gemm_probe.cu).  Forward side update for wide states with the contraction on tcgen05:
//
//   * an operator block of x1 is a K-chunk: every thread gathers its 2 x 16 bytes of block t (engine_wide.cuh
//     gathers), splits them (hi = the value itself, lo = x - trunc_tf32(x)) and stores them as two planes straight
//     into the K-major SWIZZLE_NONE core-matrix layout of a 64 x F A operand;
//   * the weight block lives in shared memory already split and laid out as the N x K (K-major) B operand of
//     every chunk;
//   * one thread issues, per chunk, F/8 x 3 tcgen05.mma.cta_group::1.kind::tf32 (lo*hi and hi*lo into two shared
//     accumulators, hi*hi into an accumulator of ITS OWN per chunk: the tensor core accumulates with truncation, so
//     no accumulator chains more than F/8 large terms; the epilogue adds the chunks in round-to-nearest fp32);
//     tcgen05.commit -> mbarrier tells the CTA when the A planes may be overwritten, so the feature loads of
//     chunk c+1 are in flight while chunk c multiplies;
//   * epilogue: warps 0-3 read their 16 rows from TMEM (M = 64: row m sits in lane 32 (m / 16) + m % 16), add bias,
//     ReLU, park the tile in shared memory; all threads store Z coalesced and keep the column statistics.
//
// Restrictions of this first version: 64-row tiles, Fs and Fc in {32, 64} (one gather item per thread), Fout in
// {32, 64}, (chunks + 2) * Fout <= 512 TMEM columns, no run-length ranges in the forward operators (none of the
// reference's operators has them), long rows gathered in line.
#pragma once

#define T5_CAP 512

__device__ __forceinline__ uint32_t t5_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle: start address, leading (K-adjacent core matrices) and stride
// (adjacent 8-row groups) byte offsets in 16-byte units, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t t5_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void t5_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// bounded: a wrong encoding traps instead of hanging the GPU
__device__ __forceinline__ void t5_wait(uint32_t mbar, uint32_t parity) {
    for (int it = 0; it < (1 << 24); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

struct Tc5Layout { int Whi, Wlo, Ahi, Alo, a_floats, bias, sc_s, sh_s, sc_c, sh_c, stage, scol, sval, pcol, pv1, pv2, total; };
__host__ __device__ inline Tc5Layout tc5_layout(int Cin, int Fout, int Fs, int Fc) {
    Tc5Layout l;
    const int Fb = Fs > Fc ? Fs : Fc;
    int o = 0;
    l.Whi = o; o += Cin * Fout;
    l.Wlo = o; o += Cin * Fout;
    l.a_floats = 64 * Fb > 64 * (Fout + 4) ? 64 * Fb : 64 * (Fout + 4);     // A plane, also the epilogue's parking tile
    l.Ahi = o; o += l.a_floats;
    l.Alo = o; o += l.a_floats;
    l.bias = o; o += Fout;
    l.sc_s = o; o += Fs;  l.sh_s = o; o += Fs;
    l.sc_c = o; o += Fc;  l.sh_c = o; o += Fc;
    o = (o + 3) & ~3;
    l.stage = o; o += (int)((sizeof(WideStage) + 15) / 16) * 4;
    l.scol = o; o += WD_SLOTS * T5_CAP;
    l.sval = o; o += WD_SLOTS * T5_CAP;
    l.pcol = o; o += T5_CAP;
    l.pv1 = o; o += T5_CAP;
    l.pv2 = o; o += T5_CAP;
    l.total = o;
    return l;
}

// float offset of element (row r of 64, column k of F) in an A plane; (output n, column k) in a chunk of a W plane
__device__ __forceinline__ int t5_a_off(int r, int k) { return (k >> 2) * 256 + (r >> 3) * 32 + (r & 7) * 4 + (k & 3); }
__device__ __forceinline__ int t5_b_off(int n, int k, int Fout) {
    return (k >> 2) * (Fout >> 3) * 32 + (n >> 3) * 32 + (n & 7) * 4 + (k & 3);
}

__global__ void __launch_bounds__(WD_THREADS, 1)
fwd_tc5_kernel(const FwdArgs a) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ double dscratch[WD_THREADS];
    __shared__ double dtot[256];
    __shared__ DeferList dl;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar_s;
    const int Cin = a.Cin, Fout = a.Fout, K = a.ops.n, Fs = a.Fs, Fc = a.Fc;
    const Tc5Layout lay = tc5_layout(Cin, Fout, Fs, Fc);
    float* Whi = smem + lay.Whi;
    float* Wlo = smem + lay.Wlo;
    float* Ahi = smem + lay.Ahi;
    float* Alo = smem + lay.Alo;
    float* bias = smem + lay.bias;
    float* sc_s = smem + lay.sc_s;
    float* sh_s = smem + lay.sh_s;
    float* sc_c = smem + lay.sc_c;
    float* sh_c = smem + lay.sh_c;
    WideStage* st = reinterpret_cast<WideStage*>(smem + lay.stage);
    int* scol = reinterpret_cast<int*>(smem + lay.scol);
    float* sval = smem + lay.sval;
    int* pcol = reinterpret_cast<int*>(smem + lay.pcol);
    float* pv1 = smem + lay.pv1;
    float* pv2 = smem + lay.pv2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool cross = a.p_rowptr != nullptr;
    const int nchunk = K + (cross ? 2 : 0);

    // ---- prologue: TMEM, mbarrier, split weight planes in B-operand order, vectors
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(t5_smem(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(t5_smem(&mbar_s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        dl.rsum_id = -1;
    }
    for (int i = tid; i < Cin * Fout; i += WD_THREADS) {
        const int o = i / Cin, c = i - o * Cin;
        const float w = (o < a.Ha) ? a.Wa[(size_t)o * Cin + c] : a.Wb[(size_t)(o - a.Ha) * Cin + c];
        int base, k;
        if (c < K * Fs) { const int t = c / Fs; k = c - t * Fs; base = t * Fs * Fout; }
        else { const int cc = c - K * Fs, j = cc / Fc; k = cc - j * Fc; base = (K * Fs + j * Fc) * Fout; }
        const int off = base + t5_b_off(o, k, Fout);
        Whi[off] = w;
        Wlo[off] = w - __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    }
    for (int o = tid; o < Fout; o += WD_THREADS)
        bias[o] = (o < a.Ha) ? (a.ba ? a.ba[o] : 0.f) : (a.bb ? a.bb[o - a.Ha] : 0.f);
    wide_assign_slots(st, a.ops);
    const bool aff_s = bn_vectors(a.bn_s, Fs, sc_s, sh_s, nullptr, nullptr, dtot, dscratch);
    const bool aff_c = cross ? bn_vectors(a.bn_c, Fc, sc_c, sh_c, nullptr, nullptr, dtot, dscratch) : false;
    const AffineLoader<4> ls{a.Xs, Fs, sc_s, sh_s, aff_s};
    const AffineLoader<4> lc{a.Xc, Fc, sc_c, sh_c, aff_c};
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");          // weight planes -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t mbar = t5_smem(&mbar_s);
    double* sstat = dtot;                      // (sum z, sum z^2) of this CTA's rows, [2 * Fout]
    for (int i = tid; i < 2 * Fout; i += WD_THREADS) sstat[i] = 0.0;

    // D = F32, A = B = TF32, both K-major, N = Fout, M = 64
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Fout >> 3) << 17) | (4u << 24);
    const uint32_t lbo_a = 8 * 128, lbo_b = (uint32_t)(Fout >> 3) * 128, sbo = 128;
    // TMEM columns: [0, Fout) lo*hi, [Fout, 2 Fout) hi*lo, then one block of Fout columns per chunk for hi*hi
    uint32_t parity = 0;
    bool pending = false;                      // MMAs that read the A planes are in flight

    // my gather items: (row, chunk pair) of the self blocks and of the cross blocks
    const int Qs = Fs >> 3, Qc = Fc >> 3;
    const int rs = tid / Qs, qs = tid - rs * Qs;
    const int rc = cross ? tid / Qc : 0, qc = cross ? tid - rc * Qc : 0;
    // column statistics: every thread owns one float4 column group of the parked tile
    const int NQ = Fout >> 2;
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
    const int ntiles = (a.R + 63) / 64;

    for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
        const int row0 = tile_id * 64;
        const int trc = min(64, a.R - row0);
        if (tid == 0) { dl.cnt = 0; dl.rng_cnt = 0; }
        __syncthreads();                       // the previous tile's parked rows have been consumed
        wide_stage<T5_CAP>(st, a.ops, a.p_rowptr, a.p_col, a.p_pm, a.p_pd, row0, trc, scol, sval, pcol, pv1, pv2, &dl);

        // one chunk: wait until the planes are free, store the split rows, hand the chunk to the tensor core
        auto put_chunk = [&](const V<4> (&v)[2], bool active, int r, int q, int F, int chunk, int wcol0) {
            if (pending) { t5_wait(mbar, parity); parity ^= 1; }
            if (active) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = (q << 2) + h * (F >> 1);
                    const int off = t5_a_off(r, k);
                    const float4 x = v[h].v;
                    float4 lo;
                    lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                    lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                    lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                    lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                    *reinterpret_cast<float4*>(Ahi + off) = x;
                    *reinterpret_cast<float4*>(Alo + off) = lo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t wbase = (uint32_t)wcol0 * (uint32_t)Fout * 4u;            // bytes into the W planes
                for (int ks = 0; ks < (F >> 3); ++ks) {
                    const uint64_t dah = t5_desc(t5_smem(Ahi) + ks * 2 * lbo_a, lbo_a, sbo);
                    const uint64_t dal = t5_desc(t5_smem(Alo) + ks * 2 * lbo_a, lbo_a, sbo);
                    const uint64_t dbh = t5_desc(t5_smem(Whi) + wbase + ks * 2 * lbo_b, lbo_b, sbo);
                    const uint64_t dbl = t5_desc(t5_smem(Wlo) + wbase + ks * 2 * lbo_b, lbo_b, sbo);
                    const uint32_t acc_small = (chunk == 0 && ks == 0) ? 0u : 1u;
                    t5_mma(tmem, dal, dbh, idesc, acc_small);
                    t5_mma(tmem + Fout, dah, dbl, idesc, acc_small);
                    t5_mma(tmem + (2 + chunk) * Fout, dah, dbh, idesc, ks == 0 ? 0u : 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
            }
            pending = true;
        };

        // ---- self blocks
        {
            const bool active = rs < 64;
            const bool valid = active && rs < trc;
            const int row = row0 + rs, xo = qs << 2, xs = Fs >> 1;
            V<4> own[2];
            own[0] = own[1] = V<4>::zero();
            if (valid) { own[0] = ls(row, xo); own[1] = ls(row, xo + xs); }
            for (int t = 0; t < K; ++t) {
                V<4> v[2];
                v[0] = v[1] = V<4>::zero();
                const int kind = a.ops.kind[t];
                if (valid) {
                    if (kind == HGNN_OP_IDENT) { v[0] = own[0]; v[1] = own[1]; }
                    else if (kind == HGNN_OP_DIAG) {
                        const float dg = __ldg(a.ops.diag[t] + row);
                        v[0] = own[0]; v[1] = own[1];
                        v[0].scale(dg); v[1].scale(dg);
                    } else {
                        const WideRow w = wide_row<T5_CAP>(st, a.ops, t, row0, rs, scol, sval);
                        wide_gather<4, 2>(w, ls, xo, xs, v);
                    }
                }
                put_chunk(v, active, rs, qs, Fs, t, t * Fs);
            }
        }
        // ---- cross blocks (Pm, Pd on one pattern: gathered together, handed over one after the other)
        if (cross) {
            const bool active = rc < 64;
            const bool valid = active && rc < trc;
            const int xo = qc << 2, xs = Fc >> 1;
            V<4> am[2], ad[2];
            am[0] = am[1] = ad[0] = ad[1] = V<4>::zero();
            if (valid) {
                const bool sm = st->staged[WD_SLOTS] != 0;
                const int base = st->base[WD_SLOTS];
                WideRow w;
                w.k0 = st->rp[WD_SLOTS][rc];
                w.k1 = st->rp[WD_SLOTS][rc + 1];
                w.smem = sm;
                w.col = sm ? pcol - base : a.p_col;
                w.val = sm ? pv1 - base : a.p_pm;
                w.val2 = sm ? pv2 - base : a.p_pd;
                wide_gather2<4, 2>(w, lc, xo, xs, am, ad);
            }
            put_chunk(am, active, rc, qc, Fc, K, K * Fs);
            put_chunk(ad, active, rc, qc, Fc, K + 1, K * Fs + Fc);
        }
        // ---- epilogue: all MMAs of the tile are done when the last commit arrives
        t5_wait(mbar, parity);
        parity ^= 1;
        pending = false;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        float* park = Ahi;                     // [64][Fout + 4]
        if (warp < 4) {
            const int m = (warp << 4) + (lane & 15);
            for (int c0 = 0; c0 < Fout; c0 += 16) {
                float sum[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] = 0.f;
                for (int blk = 0; blk < nchunk + 2; ++blk) {
                    uint32_t v[16];
                    const uint32_t addr = tmem + ((uint32_t)(warp << 5) << 16) + (uint32_t)(blk * Fout + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(addr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
                }
                if (lane < 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float z = sum[j] + bias[c0 + j];
                        if (c0 + j >= a.relu_from) z = fmaxf(z, 0.f);
                        park[m * (Fout + 4) + c0 + j] = z;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        for (int i = tid; i < 64 * NQ; i += WD_THREADS) {       // WD_THREADS is a multiple of NQ: a thread keeps its columns
            const int r = i / NQ, c4 = i - r * NQ;
            if (r < trc) {
                const float4 z = *reinterpret_cast<const float4*>(park + r * (Fout + 4) + c4 * 4);
                *reinterpret_cast<float4*>(a.Z + (size_t)(row0 + r) * Fout + c4 * 4) = z;
                st1[0] += z.x; st1[1] += z.y; st1[2] += z.z; st1[3] += z.w;
                st2[0] = fmaf(z.x, z.x, st2[0]); st2[1] = fmaf(z.y, z.y, st2[1]);
                st2[2] = fmaf(z.z, z.z, st2[2]); st2[3] = fmaf(z.w, z.w, st2[3]);
            }
        }
    }
    if (a.acc_out) {
        const int c4 = tid % NQ;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(sstat + c4 * 4 + j, (double)st1[j]);
            atomicAdd(sstat + Fout + c4 * 4 + j, (double)st2[j]);
        }
        __syncthreads();
        const int nb = hgnn_ws_bins(2 * Fout);
        for (int i = tid; i < 2 * Fout; i += WD_THREADS) accum_add(a.acc_out, 2 * Fout, nb, i, sstat[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------
// backward (EXPERIMENTAL, same status as fwd_tc5_kernel; additionally behind HGNN_B200_WIDE_TC5_BWD=1)
// ---------------------------------------------------------------------------------------------
// v1 restrictions: gPre width Fg = 64 and input width Fx = 64 (h = 32), K <= 3 gathered blocks per part (TMEM:
// 64 + 64 K + 72 K <= 512 columns), no run-length ranges (so: node sides, cross parts, GNN_simple; the edge side's
// transposed line-graph operator keeps the mma.sync kernel), long rows gathered in line.
//
// Per 64-row tile and gathered block t (a K-chunk of gX, and one dW block):
//   gX  += T_t  * Wsm_t      A = the block planes read K-major  (LBO 1024, SBO 128),  B = pre-split weight chunk
//   dW_t += T_t^T * [X | 1]   A = the SAME planes read MN-major (LBO 128, SBO 1024), B = raw input rows stored in the
//                             same plane layout plus a column of ones, so that column Fx of dW_t is sum_r T_t[r][.]:
//                             the batch-norm of the input is applied afterwards, dW = sc * dWraw + sh * colsum.
// gX: both cross terms chain in one accumulator, hi*hi gets one accumulator per block (summed in fp32 in the
// epilogue); dW: one accumulator per block for all three products, chained over all tiles of the CTA (each entry
// carries its own truncation bias, nothing sums them coherently afterwards).
#define T5B_F 64
#define T5B_NX (T5B_F + 8)         // input features + the ones column, padded to a multiple of 8

struct Tc5BwdLayout { int Whi, Wlo, Ahi, Alo, Xhi, Xlo, sc, sh, mu, rs, stage, scol, sval, total; };
__host__ __device__ inline Tc5BwdLayout tc5_bwd_layout(int nT) {
    Tc5BwdLayout l;
    int o = 0;
    l.Whi = o; o += nT * T5B_F;
    l.Wlo = o; o += nT * T5B_F;
    const int a_floats = 64 * (T5B_F + 4);             // A plane, also the epilogue's parking tile
    l.Ahi = o; o += a_floats;
    l.Alo = o; o += a_floats;
    l.Xhi = o; o += 64 * T5B_NX;
    l.Xlo = o; o += 64 * T5B_NX;
    l.sc = o; o += T5B_F;  l.sh = o; o += T5B_F;  l.mu = o; o += T5B_F;  l.rs = o; o += T5B_F;
    l.stage = o; o += (int)((sizeof(WideStage) + 15) / 16) * 4;
    l.scol = o; o += WD_SLOTS * T5_CAP;
    l.sval = o; o += WD_SLOTS * T5_CAP;
    l.total = o;
    return l;
}

__device__ __forceinline__ void bwd_tc5_part(const BwdArgs& a, const BwdPart& p, bool is_self, int first_tile,
                                             int tile_stride, float* smem, double* dscratch, double* dtot,
                                             DeferList* dl, const float* c0, const float* c1, const float* c2,
                                             bool has_bn, uint32_t tmem, uint32_t mbar, uint32_t& parity) {
    constexpr int F = T5B_F, NX = T5B_NX;
    const int K = p.ops.n, nT = K * F;
    const Tc5BwdLayout lay = tc5_bwd_layout(nT);
    float* Whi = smem + lay.Whi;               // chunk t: B operand [N = input feature f][K = gPre feature o], K-major
    float* Wlo = smem + lay.Wlo;
    float* Ahi = smem + lay.Ahi;               // gathered block: (row r, gPre feature o)
    float* Alo = smem + lay.Alo;
    float* Xhi = smem + lay.Xhi;               // raw input rows: (row r, input feature f), f = F is the ones column
    float* Xlo = smem + lay.Xlo;
    float* sc = smem + lay.sc;
    float* sh = smem + lay.sh;
    float* mu = smem + lay.mu;
    float* rs = smem + lay.rs;
    WideStage* st = reinterpret_cast<WideStage*>(smem + lay.stage);
    int* scol = reinterpret_cast<int*>(smem + lay.scol);
    float* sval = smem + lay.sval;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool want_dw = a.dW_bins != nullptr;

    __syncthreads();                           // a previous part of this CTA is done with shared memory
    if (tid == 0) dl->rsum_id = -1;
    for (int i = tid; i < nT * F; i += WD_THREADS) {
        const int c = i / F, f = i - c * F;    // c = t * F + o
        const int t = c / F, o = c - t * F;
        const float* wrow = (o < a.Ha) ? a.Wa + (size_t)o * a.Cin : a.Wb + (size_t)(o - a.Ha) * a.Cin;
        const float w = wrow[p.col0 + t * F + f];
        const int off = t * F * F + t5_b_off(f, o, F);
        Whi[off] = w;
        Wlo[off] = w - __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    }
    // the ones column (and its padding) of the input planes never changes
    for (int i = tid; i < 64 * 8; i += WD_THREADS) {
        const int r = i >> 3, f = F + (i & 7);
        Xhi[t5_a_off(r, f)] = (i & 7) == 0 ? 1.f : 0.f;
        Xlo[t5_a_off(r, f)] = 0.f;
    }
    wide_assign_slots(st, p.ops);
    bn_vectors(p.bn, F, sc, sh, mu, rs, dtot, dscratch);
    const GpreLoader<4> lg{a.gY, a.Z, F, c0, c1, c2, a.relu_from, has_bn};
    const bool dual = !is_self && K == 2 && p.ops.kind[0] == HGNN_OP_CSR && p.ops.kind[1] == HGNN_OP_CSR &&
                      p.ops.rowptr[0] == p.ops.rowptr[1] && p.ops.col[0] == p.ops.col[1];
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    double* sstat = dtot;                      // (sum g, sum g * xhat) of the rows produced here, [2 F]
    double* sdb = dtot + 2 * F;                // dbias partial sums [F]
    for (int i = tid; i < 3 * F; i += WD_THREADS) dtot[i] = 0.0;

    // gX: D = F32, A = B = TF32, K-major / K-major, N = 64, M = 64
    const uint32_t idesc_gx = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(F >> 3) << 17) | (4u << 24);
    // dW: both operands MN-major, N = 72, M = 64
    const uint32_t idesc_dw = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NX >> 3) << 17) | (4u << 24);
    // TMEM columns: [0, 64) gX cross terms, [64 (1 + t), +64) gX hi*hi of block t, [64 (1 + K) + 72 t, +72) dW of block t
    const uint32_t col_dw = 64u * (1 + K);
    bool pending = false;
    bool dw_started = false;                   // the dW accumulators hold something (first mma must not accumulate)

    const int r_ = tid >> 3, q_ = tid & 7;     // my gather item: row, chunk pair (F / 8 = 8 items per row)
    float dbacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dbacc[j] = 0.f;
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
    const int NQ = F >> 2;

    for (int tile_id = first_tile; tile_id < p.tiles; tile_id += tile_stride) {
        const int row0 = tile_id * 64;
        const int trc = min(64, p.R - row0);
        if (tid == 0) { dl->cnt = 0; dl->rng_cnt = 0; }
        __syncthreads();                       // the previous tile's parked rows / input planes have been consumed
        wide_stage<T5_CAP>(st, p.ops, nullptr, nullptr, nullptr, nullptr, row0, trc, scol, sval, nullptr, nullptr, nullptr, dl);
        // raw input rows -> planes (every mma that read them has completed: the previous epilogue waited)
        for (int i = tid; i < 64 * NQ; i += WD_THREADS) {
            const int r = i / NQ, g4 = i - r * NQ;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < trc) x = __ldg(reinterpret_cast<const float4*>(p.X + (size_t)(row0 + r) * F) + g4);
            float4 lo;
            lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            *reinterpret_cast<float4*>(Xhi + t5_a_off(r, g4 * 4)) = x;
            *reinterpret_cast<float4*>(Xlo + t5_a_off(r, g4 * 4)) = lo;
        }

        auto put_chunk = [&](const V<4> (&v)[2], int t) {
            if (pending) { t5_wait(mbar, parity); parity ^= 1; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int off = t5_a_off(r_, (q_ << 2) + h * (F >> 1));
                const float4 x = v[h].v;
                float4 lo;
                lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                *reinterpret_cast<float4*>(Ahi + off) = x;
                *reinterpret_cast<float4*>(Alo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (p.gX) {
                    const uint32_t wbase = (uint32_t)t * F * F * 4u;
                    for (int ks = 0; ks < F / 8; ++ks) {
                        const uint64_t dah = t5_desc(t5_smem(Ahi) + ks * 2048, 1024, 128);
                        const uint64_t dal = t5_desc(t5_smem(Alo) + ks * 2048, 1024, 128);
                        const uint64_t dbh = t5_desc(t5_smem(Whi) + wbase + ks * 2048, 1024, 128);
                        const uint64_t dbl = t5_desc(t5_smem(Wlo) + wbase + ks * 2048, 1024, 128);
                        const uint32_t acc_small = (t == 0 && ks == 0) ? 0u : 1u;
                        t5_mma(tmem, dal, dbh, idesc_gx, acc_small);
                        t5_mma(tmem, dah, dbl, idesc_gx, 1u);
                        t5_mma(tmem + 64u * (1 + t), dah, dbh, idesc_gx, ks == 0 ? 0u : 1u);
                    }
                }
                if (want_dw) {
                    for (int ks = 0; ks < 8; ++ks) {           // K = the 64 rows of the tile, 8 per instruction
                        const uint64_t dah = t5_desc(t5_smem(Ahi) + ks * 128, 128, 1024);
                        const uint64_t dal = t5_desc(t5_smem(Alo) + ks * 128, 128, 1024);
                        const uint64_t dbh = t5_desc(t5_smem(Xhi) + ks * 128, 128, 1024);
                        const uint64_t dbl = t5_desc(t5_smem(Xlo) + ks * 128, 128, 1024);
                        const uint32_t acc0 = (!dw_started && ks == 0) ? 0u : 1u;
                        const uint32_t d = tmem + col_dw + (uint32_t)t * NX;
                        t5_mma(d, dal, dbh, idesc_dw, acc0);
                        t5_mma(d, dah, dbl, idesc_dw, 1u);
                        t5_mma(d, dah, dbh, idesc_dw, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar) : "memory");
            }
            pending = true;
        };

        const bool valid = r_ < trc;
        const int row = row0 + r_, xo = q_ << 2, xs = F >> 1;
        if (dual) {
            V<4> am[2], ad[2];
            am[0] = am[1] = ad[0] = ad[1] = V<4>::zero();
            if (valid) {
                WideRow w = wide_row<T5_CAP>(st, p.ops, 0, row0, r_, scol, sval);
                const WideRow w1 = wide_row<T5_CAP>(st, p.ops, 1, row0, r_, scol, sval);
                w.val2 = w1.val;
                wide_gather2<3, 2>(w, lg, xo, xs, am, ad);
            }
            put_chunk(am, 0);
            put_chunk(ad, 1);
        } else {
            V<4> own[2];
            own[0] = own[1] = V<4>::zero();
            if (valid) { own[0] = lg(row, xo); own[1] = lg(row, xo + xs); }
            if (is_self) {                     // dbias = column sums of the own gPre rows
                dbacc[0] += own[0].v.x; dbacc[1] += own[0].v.y; dbacc[2] += own[0].v.z; dbacc[3] += own[0].v.w;
                dbacc[4] += own[1].v.x; dbacc[5] += own[1].v.y; dbacc[6] += own[1].v.z; dbacc[7] += own[1].v.w;
            }
            for (int t = 0; t < K; ++t) {
                V<4> v[2];
                v[0] = v[1] = V<4>::zero();
                const int kind = p.ops.kind[t];
                if (valid) {
                    if (kind == HGNN_OP_IDENT) { v[0] = own[0]; v[1] = own[1]; }
                    else if (kind == HGNN_OP_DIAG) {
                        const float dg = __ldg(p.ops.diag[t] + row);
                        v[0] = own[0]; v[1] = own[1];
                        v[0].scale(dg); v[1].scale(dg);
                    } else {
                        const WideRow w = wide_row<T5_CAP>(st, p.ops, t, row0, r_, scol, sval);
                        wide_gather<3, 2>(w, lg, xo, xs, v);
                    }
                }
                put_chunk(v, t);
            }
        }
        dw_started = dw_started || want_dw;
        // ---- all mma of the tile done: gX epilogue
        t5_wait(mbar, parity);
        parity ^= 1;
        pending = false;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (p.gX) {
            float* park = Ahi;                 // [64][F + 4]
            if (warp < 4) {
                const int m = (warp << 4) + (lane & 15);
                for (int c0_ = 0; c0_ < F; c0_ += 16) {
                    float sum[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[j] = 0.f;
                    for (int blk = 0; blk < 1 + K; ++blk) {
                        uint32_t v[16];
                        const uint32_t addr = tmem + ((uint32_t)(warp << 5) << 16) + (uint32_t)(blk * 64 + c0_);
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                            : "r"(addr) : "memory");
                        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[j] += __uint_as_float(v[j]);
                    }
                    if (lane < 16) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) park[m * (F + 4) + c0_ + j] = sum[j];
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            for (int i = tid; i < 64 * NQ; i += WD_THREADS) {      // a thread keeps its column group (512 = 32 x NQ)
                const int r = i / NQ, c4 = i - r * NQ;
                if (r < trc) {
                    float4 gx = *reinterpret_cast<const float4*>(park + r * (F + 4) + c4 * 4);
                    if (p.acc_b) {
                        const float4 x = *reinterpret_cast<const float4*>(Xhi + t5_a_off(r, c4 * 4));
                        const int f = c4 * 4;
                        st1[0] += gx.x; st1[1] += gx.y; st1[2] += gx.z; st1[3] += gx.w;
                        st2[0] = fmaf(gx.x, (x.x - mu[f]) * rs[f], st2[0]);
                        st2[1] = fmaf(gx.y, (x.y - mu[f + 1]) * rs[f + 1], st2[1]);
                        st2[2] = fmaf(gx.z, (x.z - mu[f + 2]) * rs[f + 2], st2[2]);
                        st2[3] = fmaf(gx.w, (x.w - mu[f + 3]) * rs[f + 3], st2[3]);
                    }
                    float4* dst = reinterpret_cast<float4*>(p.gX + (size_t)(row0 + r) * F + c4 * 4);
                    if (p.accumulate) { const float4 old = *dst; gx.x += old.x; gx.y += old.y; gx.z += old.z; gx.w += old.w; }
                    *dst = gx;
                }
            }
        }
    }
    // ---- end of the part: dW from TMEM, dbias, statistics
    __syncthreads();
    if (want_dw && dw_started) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int nbw = hgnn_ws_bins(F * a.Cin);
        if (warp < 4) {
            const int o = (warp << 4) + (lane & 15);
            for (int t = 0; t < K; ++t) {
                const uint32_t base = tmem + ((uint32_t)(warp << 5) << 16) + col_dw + (uint32_t)t * NX;
                uint32_t cs[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                             : "=r"(cs[0]), "=r"(cs[1]), "=r"(cs[2]), "=r"(cs[3]), "=r"(cs[4]), "=r"(cs[5]), "=r"(cs[6]), "=r"(cs[7])
                             : "r"(base + F) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                const float colsum = __uint_as_float(cs[0]);       // sum_r T_t[r][o]
                for (int f0 = 0; f0 < F; f0 += 8) {
                    uint32_t v[8];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                 : "r"(base + f0) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    if (lane < 16) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int f = f0 + j;
                            const float g = fmaf(sc[f], __uint_as_float(v[j]), sh[f] * colsum);
                            accum_add(a.dW_bins, F * a.Cin, nbw, o * a.Cin + p.col0 + t * F + f, (double)g);
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        if (is_self && a.db_bins) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(sdb + (q_ << 2) + (j & 3) + (j >> 2) * (F >> 1), (double)dbacc[j]);
            __syncthreads();
            const int nbb = hgnn_ws_bins(F);
            for (int o = tid; o < F; o += WD_THREADS) accum_add(a.db_bins, F, nbb, o, sdb[o]);
        }
    }
    if (p.acc_b && p.gX) {
        const int c4 = tid % NQ;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(sstat + c4 * 4 + j, (double)st1[j]);
            atomicAdd(sstat + F + c4 * 4 + j, (double)st2[j]);
        }
        __syncthreads();
        const int nb = hgnn_ws_bins(2 * F);
        for (int i = tid; i < 2 * F; i += WD_THREADS) accum_add(p.acc_b, 2 * F, nb, i, sstat[i]);
    }
}

__global__ void __launch_bounds__(WD_THREADS, 1)
bwd_tc5_kernel(const BwdArgs a) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ double dscratch[WD_THREADS];
    __shared__ double dtot[512];
    __shared__ DeferList dl;
    __shared__ __align__(16) float coef[3 * 128];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar_s;
    const int Fg = a.Fg;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(t5_smem(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(t5_smem(&mbar_s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        dl.rsum_id = -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    float* c0 = coef;
    float* c1 = coef + 128;
    float* c2 = coef + 256;
    const bool has_bn = a.acc_b != nullptr;
    if (has_bn) {
        // coefficients of the BN backward of THIS side: gZ = c0 g + c1 + c2 z  (batch_normalization.py:65-77)
        double* tf = dtot;
        double* tb = dtot + 2 * Fg;
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
    uint32_t parity = 0;
    const uint32_t mbar = t5_smem(&mbar_s);
    if (a.self.R > 0)
        bwd_tc5_part(a, a.self, true, blockIdx.x, gridDim.x, smem, dscratch, dtot, &dl, c0, c1, c2, has_bn, tmem, mbar, parity);
    if (a.cross.R > 0)
        bwd_tc5_part(a, a.cross, false, gridDim.x - 1 - blockIdx.x, gridDim.x, smem, dscratch, dtot, &dl, c0, c1, c2, has_bn,
                     tmem, mbar, parity);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}
