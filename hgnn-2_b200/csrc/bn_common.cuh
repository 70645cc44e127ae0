// bn_common.cuh -- shared pieces of the masked batch-norm epilogue.
// Reference: models/layers/batch_normalization.py:34-43 (BN.forward), :65-77 (sb_normalization),
// :80-93 (mean_with_padding).  In the packed layout every row is a real slot, so the padding mask
// disappears and n = R.
#pragma once
#include "common.cuh"

#define HGNN_BN_EPS 1e-5

// Column-owner mapping used by every kernel that reduces per-feature sums over rows:
// thread t owns feature t % F and walks rows t / F, t / F + rows_per_pass, ...
struct ColOwner {
    int f, rg, rows_per_pass;
    bool active;
    __device__ ColOwner(int F, int nthreads) {
        rows_per_pass = nthreads / F;
        active = (int)threadIdx.x < rows_per_pass * F;
        f = threadIdx.x % F;
        rg = threadIdx.x / F;
    }
};

// Reduce per-thread (s1, s2) of column owners across the CTA (fixed order), then add the CTA's sums
// into the fp64 accumulators accum[0..F) (s1) and accum[F..2F) (s2).  `red` = 2*blockDim doubles.
__device__ __forceinline__ void cta_column_accumulate(double s1, double s2, int F, const ColOwner& co,
                                                      double* red, double* accum) {
    double a = 0.0, b = 0.0;
    if ((32 % F) == 0 && (blockDim.x % F) == 0) {      // every thread is an owner: shuffle tree
        a = cta_reduce_mod(s1, F, red);
        b = cta_reduce_mod(s2, F, red);
    } else {
        red[threadIdx.x] = co.active ? s1 : 0.0;
        red[blockDim.x + threadIdx.x] = co.active ? s2 : 0.0;
        __syncthreads();
        if ((int)threadIdx.x < F) {
            for (int k = 0; k < co.rows_per_pass; ++k) {
                a += red[k * F + threadIdx.x];
                b += red[blockDim.x + k * F + threadIdx.x];
            }
        }
    }
    if ((int)threadIdx.x < F) {
        const int nb = hgnn_ws_bins(2 * F);
        accum_add(accum, 2 * F, nb, threadIdx.x, a);
        accum_add(accum, 2 * F, nb, F + threadIdx.x, b);
    }
}

// Last CTA: read (and re-zero) the accumulated sums, emit stats = [mean, std, scale, shift] (4F).
__device__ __forceinline__ void bn_finalize_accum(double* accum, int F, long long n,
                                                  const float* weight, const float* bias,
                                                  float* running_mean, float* running_std,
                                                  float momentum, float* stats) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const int nb = hgnn_ws_bins(2 * F);
        const double a = accum_take(accum, 2 * F, nb, f), b = accum_take(accum, 2 * F, nb, F + f);
        double mean = a / (double)n;
        double var = b / (double)n - mean * mean;
        if (var < 0.0) var = 0.0;
        double sd = sqrt(var + HGNN_BN_EPS);
        float w = weight ? weight[0] : 1.f, bb = bias ? bias[0] : 0.f;
        stats[f] = (float)mean;
        stats[F + f] = (float)sd;
        stats[2 * F + f] = (float)((double)w / sd);
        stats[3 * F + f] = (float)((double)bb - (double)w * mean / sd);
        if (running_mean) running_mean[f] = (1.f - momentum) * (float)mean + momentum * running_mean[f];
        if (running_std) running_std[f] = (1.f - momentum) * (float)sd + momentum * running_std[f];
    }
}
