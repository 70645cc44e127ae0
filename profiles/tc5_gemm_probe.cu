// tc5_gemm_probe.cu -- bring-up probe for the planned tcgen05 version of the wide contractions.
// NOT part of the product library; a stand-alone program:
//     nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tc5_probe profiles/tc5_gemm_probe.cu
//     timeout 30 gpurun_out/tc5_probe
// One CTA computes D[64 x 64] = A[64 x K] * B[64 x K]^T (K = 64 per chunk, NCHUNK chunks) with the 3xTF32 split
// (hi = the fp32 value itself - the tensor core ignores the low 13 mantissa bits -, lo = x - trunc_tf32(x)) through
// tcgen05.mma.cta_group::1.kind::tf32, operands written by the threads straight into the K-major SWIZZLE_NONE
// core-matrix layout (what the gather of engine_wide.cuh would do), accumulator in TMEM, read back with
// tcgen05.ld.32x32b.  The host checks against fp64 and reports
//   * which of the two readings of (leading, stride) byte offsets of the shared-memory descriptor is right,
//   * which TMEM lane holds which output row for M = 64,
//   * the error of one chained accumulator against separate accumulators for the three products.
// Every wait is bounded: a wrong encoding prints "TIMEOUT", it does not hang the GPU.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define TM 64
#define TN 64
#define KC 64                  // K per chunk (one operator block at h = 32)
#define PLANE (TM * KC)        // floats per operand plane of a chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// element (row r, column k) of a 64 x 64 K-major SWIZZLE_NONE operand: core matrix = 8 rows x 16 bytes (4 tf32),
// core matrices adjacent in K are `lbo` bytes apart, adjacent 8-row groups `sbo` bytes apart
__device__ __forceinline__ int core_offset_floats(int r, int k, int lbo, int sbo) {
    return ((r >> 3) * sbo + (k >> 2) * lbo + (r & 7) * 16 + (k & 3) * 4) >> 2;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo, int swap) {
    const uint64_t a = (saddr >> 4) & 0x3fff;
    const uint64_t l = ((swap ? sbo : lbo) >> 4) & 0x3fff;
    const uint64_t s = ((swap ? lbo : sbo) >> 4) & 0x3fff;
    // bits [46,48): descriptor version = 1 on sm_100; layout type (bits 61-63) = 0: SWIZZLE_NONE
    return a | (l << 16) | (s << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

// mode bit 0: swap the two byte-offset fields; bit 1: three accumulators (lo*hi, hi*lo, hi*hi) instead of one
__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ B, int nchunk, int mode, float* __restrict__ out,
             int* __restrict__ status) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar_s;
    float* a_hi = smem;
    float* a_lo = a_hi + PLANE;
    float* b_hi = a_lo + PLANE;
    float* b_lo = b_hi + PLANE;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int lbo = 1024, sbo = 128;       // K-adjacent core matrices 1024 B apart (8 row groups x 128 B), row groups 128 B
    const int K = nchunk * KC;
    const bool three = (mode & 2) != 0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar_s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
    // N >> 3 at bits 17-22, M >> 4 at bits 24-28
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    uint32_t parity = 0;
    bool ok = true;

    for (int c = 0; c < nchunk; ++c) {
        // the "gather": split and store one 64 x 64 chunk of each operand in core-matrix order
        for (int i = tid; i < TM * KC; i += 128) {
            const int r = i / KC, k = i - r * KC;
            const float av = A[(size_t)r * K + c * KC + k], bv = B[(size_t)r * K + c * KC + k];
            const int o = core_offset_floats(r, k, lbo, sbo);
            a_hi[o] = av;
            a_lo[o] = av - __uint_as_float(__float_as_uint(av) & 0xffffe000u);
            b_hi[o] = bv;
            b_lo[o] = bv - __uint_as_float(__float_as_uint(bv) & 0xffffe000u);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy stores -> async proxy (UMMA)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int swap = mode & 1;
            for (int ks = 0; ks < KC / 8; ++ks) {          // one instruction = K of 8 tf32 = two core matrices in K
                const uint32_t koff = ks * 2 * lbo;
                const uint64_t dah = make_desc(smem_u32(a_hi) + koff, lbo, sbo, swap);
                const uint64_t dal = make_desc(smem_u32(a_lo) + koff, lbo, sbo, swap);
                const uint64_t dbh = make_desc(smem_u32(b_hi) + koff, lbo, sbo, swap);
                const uint64_t dbl = make_desc(smem_u32(b_lo) + koff, lbo, sbo, swap);
                const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
                if (three) {
                    mma_tf32(tmem + 0 * TN, dal, dbh, idesc, first);
                    mma_tf32(tmem + 1 * TN, dah, dbl, idesc, first);
                    mma_tf32(tmem + 2 * TN, dah, dbh, idesc, first);
                } else {
                    mma_tf32(tmem, dal, dbh, idesc, first);
                    mma_tf32(tmem, dah, dbl, idesc, 1u);
                    mma_tf32(tmem, dah, dbh, idesc, 1u);
                }
            }
            // arrives on the mbarrier once every mma issued so far has read its operands and written TMEM
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar_s)) : "memory");
        }
        ok = mbar_wait(smem_u32(&mbar_s), parity) && ok;   // also guards the reuse of the operand planes
        parity ^= 1;
        __syncthreads();
        if (!ok) break;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0) status[0] = ok ? 1 : -1;
    if (ok) {
        // every warp reads its 32 TMEM lanes: thread = lane, 64 columns per accumulator
        const int nacc = three ? 3 : 1;
        for (int acc = 0; acc < nacc; ++acc) {
            for (int c0 = 0; c0 < TN; c0 += 8) {
                uint32_t v[8];
                const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + acc * TN + c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                             : "r"(addr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                for (int j = 0; j < 8; ++j) out[((size_t)acc * 128 + tid) * TN + c0 + j] = __uint_as_float(v[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

// Second question for the backward: the SAME planes read MN-major.  E[f1][f2] = sum_r A[r][f1] B[r][f2]
// over the 64 rows of chunk 0 (the dW contraction: M = feature of A, N = feature of B, K = row), a_major = b_major = 1,
// k-groups of 8 rows 128 B apart (leading), m-groups of 4 features 1024 B apart (stride); mode bit 0 swaps the two.
__global__ void __launch_bounds__(128, 1)
probe_mn_kernel(const float* __restrict__ A, const float* __restrict__ B, int K, int mode, int nx, float* __restrict__ out,
                int* __restrict__ status) {
    extern __shared__ __align__(1024) float smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar_s;
    float* a_hi = smem;
    float* a_lo = a_hi + PLANE;
    float* b_hi = a_lo + PLANE;
    float* b_lo = b_hi + TM * 72;              // room for nx = 72: 64 features + a column of ones + padding
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < TM * 8; i += 128) {  // columns 64..71 of the B planes (only read when nx = 72)
        const int r = i >> 3, f = 64 + (i & 7);
        b_hi[core_offset_floats(r, f, 1024, 128)] = (i & 7) == 0 ? 1.f : 0.f;
        b_lo[core_offset_floats(r, f, 1024, 128)] = 0.f;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar_s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < TM * KC; i += 128) {
        const int r = i / KC, k = i - r * KC;
        const float av = A[(size_t)r * K + k], bv = B[(size_t)r * K + k];
        const int o = core_offset_floats(r, k, 1024, 128);
        a_hi[o] = av;
        a_lo[o] = av - __uint_as_float(__float_as_uint(av) & 0xffffe000u);
        b_hi[o] = bv;
        b_lo[o] = bv - __uint_as_float(__float_as_uint(bv) & 0xffffe000u);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(nx >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    if (tid == 0) {
        const int lbo = 128, sbo = 1024, swap = mode & 1;
        for (int ks = 0; ks < TM / 8; ++ks) {              // K = 8 rows per instruction: one k-group
            const uint32_t koff = ks * 128;
            const uint64_t dah = make_desc(smem_u32(a_hi) + koff, lbo, sbo, swap);
            const uint64_t dal = make_desc(smem_u32(a_lo) + koff, lbo, sbo, swap);
            const uint64_t dbh = make_desc(smem_u32(b_hi) + koff, lbo, sbo, swap);
            const uint64_t dbl = make_desc(smem_u32(b_lo) + koff, lbo, sbo, swap);
            mma_tf32(tmem, dal, dbh, idesc, ks == 0 ? 0u : 1u);
            mma_tf32(tmem, dah, dbl, idesc, 1u);
            mma_tf32(tmem, dah, dbh, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar_s)) : "memory");
    }
    const bool ok = mbar_wait(smem_u32(&mbar_s), 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0) status[0] = ok ? 1 : -1;
    if (ok) {
        for (int c0 = 0; c0 < nx; c0 += 8) {
            uint32_t v[8];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(addr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 8; ++j) out[(size_t)tid * 72 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem) : "memory");
}

int main() {
    const int nchunk = 5, K = nchunk * KC;
    std::vector<float> A(TM * K), B(TN * K);
    srand(1);
    for (auto& x : A) x = (float)((rand() / (double)RAND_MAX - 0.5) * exp(3.0 * (rand() / (double)RAND_MAX - 0.5)));
    for (auto& x : B) x = (float)((rand() / (double)RAND_MAX - 0.5) * 0.2);
    std::vector<double> ref(TM * TN), scale(TM * TN);
    for (int m = 0; m < TM; ++m)
        for (int n = 0; n < TN; ++n) {
            double s = 0, a = 0;
            for (int k = 0; k < K; ++k) { s += (double)A[m * K + k] * B[n * K + k]; a += fabs((double)A[m * K + k] * B[n * K + k]); }
            ref[m * TN + n] = s; scale[m * TN + n] = a;
        }
    float *dA, *dB, *dOut; int* dStatus;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dOut, 3 * 128 * TN * 4); cudaMalloc(&dStatus, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 4 * PLANE * sizeof(float);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> out(3 * 128 * TN);
    for (int mode = 0; mode < 4; ++mode) {
        cudaMemset(dOut, 0, out.size() * 4); cudaMemset(dStatus, 0, 4);
        probe_kernel<<<1, 128, smem>>>(dA, dB, nchunk, mode, dOut, dStatus);
        cudaError_t e = cudaDeviceSynchronize();
        int status = 0;
        cudaMemcpy(&status, dStatus, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        printf("mode %d (swap offsets %d, three accumulators %d): %s, status %d%s\n", mode, mode & 1, (mode >> 1) & 1,
               cudaGetErrorString(e), status, status < 0 ? "  TIMEOUT waiting for tcgen05.commit" : "");
        if (e != cudaSuccess || status <= 0) { if (e != cudaSuccess) return 1; continue; }
        const int nacc = (mode & 2) ? 3 : 1;
        // which TMEM lane holds output row m?  best-matching lane per row
        int lane_of[TM]; double worst = 0, sum_signed = 0, sum_abs = 0;
        for (int m = 0; m < TM; ++m) {
            double best = 1e300; int bl = -1;
            for (int l = 0; l < 128; ++l) {
                double err = 0;
                for (int n = 0; n < TN; ++n) {
                    double v = 0;
                    for (int a = 0; a < nacc; ++a) v += out[((size_t)a * 128 + l) * TN + n];
                    err = fmax(err, fabs(v - ref[m * TN + n]) / scale[m * TN + n]);
                }
                if (err < best) { best = err; bl = l; }
            }
            lane_of[m] = bl; worst = fmax(worst, best);
            for (int n = 0; n < TN; ++n) {
                double v = 0;
                for (int a = 0; a < nacc; ++a) v += out[((size_t)a * 128 + bl) * TN + n];
                sum_signed += v - ref[m * TN + n]; sum_abs += fabs(ref[m * TN + n]);
            }
        }
        printf("  max |err| / sum|a||b| = %.3e   signed error sum / sum|ref| = %.3e   (fp32-grade: < 2e-6 / < 1e-5)\n",
               worst, sum_signed / sum_abs);
        printf("  row -> TMEM lane:");
        for (int m = 0; m < TM; m += 8) printf(" %d:%d", m, lane_of[m]);
        printf("\n");
    }
    // ---- MN-major reading of the same planes (the dW contraction of the backward)
    std::vector<double> eref(TM * TN), escale(TM * TN);
    for (int f1 = 0; f1 < TM; ++f1)
        for (int f2 = 0; f2 < TN; ++f2) {
            double s2 = 0, a2 = 0;
            for (int r = 0; r < TM; ++r) { s2 += (double)A[r * K + f1] * B[r * K + f2]; a2 += fabs((double)A[r * K + f1] * B[r * K + f2]); }
            eref[f1 * TN + f2] = s2; escale[f1 * TN + f2] = a2;
        }
    const size_t smem_mn = (2 * PLANE + 2 * TM * 72) * sizeof(float);
    cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mn);
    for (int mode = 0; mode < 4; ++mode) {         // bit 0: swap the offsets; bit 1: N = 72 (ones column: the dW operand)
        const int nx = (mode & 2) ? 72 : 64;
        cudaMemset(dOut, 0, out.size() * 4); cudaMemset(dStatus, 0, 4);
        probe_mn_kernel<<<1, 128, smem_mn>>>(dA, dB, K, mode, nx, dOut, dStatus);
        cudaError_t e = cudaDeviceSynchronize();
        int status = 0;
        cudaMemcpy(&status, dStatus, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), dOut, 128 * 72 * 4, cudaMemcpyDeviceToHost);
        printf("MN-major mode %d (swap offsets %d, N = %d): %s, status %d\n", mode, mode & 1, nx, cudaGetErrorString(e), status);
        if (e != cudaSuccess) return 1;
        if (status <= 0) continue;
        double worst = 0;
        for (int m = 0; m < TM; ++m) {
            const int l = (m / 16) * 32 + m % 16;          // M = 64: half sub-partitions
            for (int n = 0; n < TN; ++n) worst = fmax(worst, fabs(out[(size_t)l * 72 + n] - eref[m * TN + n]) / escale[m * TN + n]);
        }
        printf("  max |err| / sum|a||b| = %.3e (rows taken from lanes 32 (m / 16) + m %% 16)\n", worst);
        if (nx == 72) {                            // column 64 = sum over the 64 rows of A[r][m]
            double wc = 0;
            for (int m = 0; m < TM; ++m) {
                const int l = (m / 16) * 32 + m % 16;
                double cs = 0, ca = 0;
                for (int r = 0; r < TM; ++r) { cs += A[r * K + m]; ca += fabs(A[r * K + m]); }
                wc = fmax(wc, fabs(out[(size_t)l * 72 + 64] - cs) / ca);
            }
            printf("  ones column: max |colsum err| / sum|a| = %.3e\n", wc);
        }
    }
    return 0;
}
