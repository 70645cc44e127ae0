// bn.cu -- masked batch-norm on packed rows: statistics, apply, backward.
// Reference: models/layers/batch_normalization.py:23-108.  HBM-bound elementwise / column
// reductions; cross-CTA sums go through fp64 atomics into the workspace accumulators and the last
// CTA (ticket) finalises them in the same launch.
#include "bn_common.cuh"

int hgnn_grid_cap(int width);

#define BN_MAX_F 128

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ Z, int R, int F, const float* weight, const float* bias,
                float* running_mean, float* running_std, float momentum, float* stats,
                unsigned int* counter, double* partial) {
    __shared__ double red[768];
    ColOwner co(F, blockDim.x);
    double s1 = 0.0, s2 = 0.0;
    if (co.active) {
        for (long long r = (long long)blockIdx.x * co.rows_per_pass + co.rg; r < R;
             r += (long long)gridDim.x * co.rows_per_pass) {
            double x = (double)Z[r * F + co.f];
            s1 += x;
            s2 += x * x;
        }
    }
    cta_column_accumulate(s1, s2, F, co, red, partial);
    if (last_block_ticket(counter)) {
        bn_finalize_accum(partial, F, R, weight, bias, running_mean, running_std, momentum, stats);
        if (threadIdx.x == 0) *counter = 0;
    }
}

extern "C" int hgnn_bn_stats(const float* Z, int R, int F, const float* weight, const float* bias,
                             float* running_mean, float* running_std, float momentum, float* stats,
                             void* ws, long long ws_bytes, hgnn_stream_t stream) {
    HGNN_REQUIRE(Z && stats && ws && R > 0, "bad argument");
    HGNN_REQUIRE(F >= 1 && F <= BN_MAX_F, "feature width must be in [1, 128]");
    if (ws_bytes < hgnn_workspace_bytes(2 * F)) {
        hgnn_set_error("hgnn_bn_stats: workspace too small");
        return HGNN_ERR_WORKSPACE;
    }
    int rows_per_cta = 256 / F;
    int grid = min(ceil_div(R, rows_per_cta * 4), HGNN_SM_COUNT * 8);
    bn_stats_kernel<<<grid, 256, 0, to_stream(stream)>>>(
        Z, R, F, weight, bias, running_mean, running_std, momentum, stats, (unsigned int*)ws,
        (double*)((char*)ws + HGNN_WS_HEADER));
    return hgnn_check_launch("hgnn_bn_stats");
}

__global__ void bn_stats_eval_kernel(int F, const float* weight, const float* bias,
                                     const float* running_mean, const float* running_std,
                                     float* stats) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float m = running_mean[f], s = running_std[f];
        float w = weight ? weight[0] : 1.f, b = bias ? bias[0] : 0.f;
        stats[f] = m;
        stats[F + f] = s;
        stats[2 * F + f] = w / s;
        stats[3 * F + f] = b - w * m / s;
    }
}

extern "C" int hgnn_bn_stats_eval(int F, const float* weight, const float* bias,
                                  const float* running_mean, const float* running_std, float* stats,
                                  hgnn_stream_t stream) {
    HGNN_REQUIRE(F >= 1 && running_mean && running_std && stats, "bad argument");
    bn_stats_eval_kernel<<<1, 256, 0, to_stream(stream)>>>(F, weight, bias, running_mean, running_std, stats);
    return hgnn_check_launch("hgnn_bn_stats_eval");
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ Z, long long n, int F, const float* __restrict__ stats,
                float* __restrict__ Y) {
    extern __shared__ float ss[];  // scale[F], shift[F]
    for (int f = threadIdx.x; f < 2 * F; f += blockDim.x) ss[f] = stats[2 * F + f];
    __syncthreads();
    if ((F & 3) == 0) {
        const float4* z4 = reinterpret_cast<const float4*>(Z);
        float4* y4 = reinterpret_cast<float4*>(Y);
        const long long n4 = n >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
             i += (long long)gridDim.x * blockDim.x) {
            float4 v = z4[i];
            int f = (int)((i * 4) % F);
            v.x = v.x * ss[f] + ss[F + f];
            v.y = v.y * ss[f + 1] + ss[F + f + 1];
            v.z = v.z * ss[f + 2] + ss[F + f + 2];
            v.w = v.w * ss[f + 3] + ss[F + f + 3];
            y4[i] = v;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (long long)gridDim.x * blockDim.x) {
            int f = (int)(i % F);
            Y[i] = Z[i] * ss[f] + ss[F + f];
        }
    }
}

extern "C" int hgnn_bn_apply(const float* Z, int R, int F, const float* stats, float* Y,
                             hgnn_stream_t stream) {
    HGNN_REQUIRE(R >= 0 && F >= 1 && stats, "bad argument");
    if (R == 0) return HGNN_OK;
    long long n = (long long)R * F;
    int grid = persistent_grid(ceil_div(n, 256 * 4), 8);
    bn_apply_kernel<<<grid, 256, 2 * F * sizeof(float), to_stream(stream)>>>(Z, n, F, stats, Y);
    return hgnn_check_launch("hgnn_bn_apply");
}

// ---------------------------------------------------------------------------------------------
// backward step 1: per-feature sums of g and g*xhat -> affine coefficients of the input gradient
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ gY, const float* __restrict__ Z, int R, int F,
                     const float* __restrict__ stats, const float* weight, int train,
                     const float* __restrict__ gshift, float* coef, unsigned int* counter,
                     double* partial) {
    __shared__ double red[768];
    ColOwner co(F, blockDim.x);
    double s1 = 0.0, s2 = 0.0;
    if (co.active) {
        const float mean = stats[co.f], rstd = 1.f / stats[F + co.f];
        for (long long r = (long long)blockIdx.x * co.rows_per_pass + co.rg; r < R;
             r += (long long)gridDim.x * co.rows_per_pass) {
            float g = gY[r * F + co.f];
            float xh = (Z[r * F + co.f] - mean) * rstd;
            s1 += (double)g;
            s2 += (double)g * (double)xh;
        }
    }
    cta_column_accumulate(s1, s2, F, co, red, partial);
    if (last_block_ticket(counter)) {
        double* tot = red + 512;
        double a_keep = 0.0, b_keep = 0.0;
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            const int nb = hgnn_ws_bins(2 * F);
            double a = accum_take(partial, 2 * F, nb, f), b = accum_take(partial, 2 * F, nb, F + f);
            double mean = stats[f], sd = stats[F + f];
            double c0 = (double)weight[0] / sd;
            double c1 = 0.0, c2 = 0.0;
            if (train) {
                c2 = -c0 * b / ((double)R * sd);
                c1 = -c0 * a / (double)R - c2 * mean;
            }
            // gradient arriving through the padded slots' fill value shift = bias - weight*mean/std
            // (batch_normalization.py:75 normalises padded slots too): affine in z as well.
            if (gshift) {
                const double sft = gshift[f], w = weight[0];
                a += sft;                       // d bias
                b += -sft * mean / sd;          // d weight
                if (train) {
                    const double k = sft * w * mean / ((double)R * sd * sd * sd);
                    c2 += k;
                    c1 += -sft * w / ((double)R * sd) - k * mean;
                }
            }
            a_keep = a;
            b_keep = b;
            coef[f] = (float)c0;
            coef[F + f] = (float)c1;
            coef[2 * F + f] = (float)c2;
        }
        __syncthreads();
        if ((int)threadIdx.x < F) {      // F <= 256 = blockDim: one feature per thread
            tot[threadIdx.x] = a_keep;
            tot[F + threadIdx.x] = b_keep;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double gw = 0.0, gb = 0.0;
            for (int f = 0; f < F; ++f) {
                gb += tot[f];
                gw += tot[F + f];
            }
            coef[3 * F] = (float)gw;
            coef[3 * F + 1] = (float)gb;
            *counter = 0;
        }
    }
}

extern "C" int hgnn_bn_bwd_reduce(const float* gY, const float* Z, int R, int F, const float* stats,
                                  const float* weight, int train, const float* gshift, float* coef,
                                  void* ws, long long ws_bytes, hgnn_stream_t stream) {
    HGNN_REQUIRE(gY && Z && stats && weight && coef && ws && R > 0, "bad argument");
    HGNN_REQUIRE(F >= 1 && F <= BN_MAX_F, "feature width must be in [1, 128]");
    if (ws_bytes < hgnn_workspace_bytes(2 * F)) {
        hgnn_set_error("hgnn_bn_bwd_reduce: workspace too small");
        return HGNN_ERR_WORKSPACE;
    }
    int rows_per_cta = 256 / F;
    int grid = min(ceil_div(R, rows_per_cta * 4), HGNN_SM_COUNT * 8);
    bn_bwd_reduce_kernel<<<grid, 256, 0, to_stream(stream)>>>(
        gY, Z, R, F, stats, weight, train, gshift, coef, (unsigned int*)ws,
        (double*)((char*)ws + HGNN_WS_HEADER));
    return hgnn_check_launch("hgnn_bn_bwd_reduce");
}

// ---------------------------------------------------------------------------------------------
// backward step 2: gPre = (c0*g + c1 + c2*Z) * relu_mask ; dbias = column sums of gPre
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
side_bwd_pre_kernel(const float* __restrict__ gY, const float* __restrict__ Z, int R, int F,
                    const float* __restrict__ coef, int relu_from, float* __restrict__ gPre,
                    float* dbias, unsigned int* counter, double* partial) {
    __shared__ double red[768];
    ColOwner co(F, blockDim.x);
    double s1 = 0.0;
    if (co.active) {
        float c0 = 1.f, c1 = 0.f, c2 = 0.f;
        if (coef) {
            c0 = coef[co.f];
            c1 = coef[F + co.f];
            c2 = coef[2 * F + co.f];
        }
        const bool relu = co.f >= relu_from;
        for (long long r = (long long)blockIdx.x * co.rows_per_pass + co.rg; r < R;
             r += (long long)gridDim.x * co.rows_per_pass) {
            float z = Z[r * F + co.f];
            float g = c0 * gY[r * F + co.f] + c1 + c2 * z;
            if (relu && !(z > 0.f)) g = 0.f;
            gPre[r * F + co.f] = g;
            s1 += (double)g;
        }
    }
    cta_column_accumulate(s1, 0.0, F, co, red, partial);
    if (last_block_ticket(counter)) {
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            const int nb = hgnn_ws_bins(2 * F);
            const double a = accum_take(partial, 2 * F, nb, f);
            accum_take(partial, 2 * F, nb, F + f);
            if (dbias) dbias[f] = (float)a;
        }
        if (threadIdx.x == 0) *counter = 0;
    }
}

extern "C" int hgnn_side_bwd_pre(const float* gY, const float* Z, int R, int F, const float* coef,
                                 int relu_from, float* gPre, float* dbias, void* ws,
                                 long long ws_bytes, hgnn_stream_t stream) {
    HGNN_REQUIRE(gY && Z && gPre && ws && R > 0, "bad argument");
    HGNN_REQUIRE(F >= 1 && F <= BN_MAX_F, "feature width must be in [1, 128]");
    if (ws_bytes < hgnn_workspace_bytes(2 * F)) {
        hgnn_set_error("hgnn_side_bwd_pre: workspace too small");
        return HGNN_ERR_WORKSPACE;
    }
    int rows_per_cta = 256 / F;
    int grid = min(ceil_div(R, rows_per_cta * 4), HGNN_SM_COUNT * 8);
    side_bwd_pre_kernel<<<grid, 256, 0, to_stream(stream)>>>(
        gY, Z, R, F, coef, relu_from, gPre, dbias, (unsigned int*)ws,
        (double*)((char*)ws + HGNN_WS_HEADER));
    return hgnn_check_launch("hgnn_side_bwd_pre");
}
