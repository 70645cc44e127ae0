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
