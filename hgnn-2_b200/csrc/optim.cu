// optim.cu -- fused Adamax over one flat fp32 parameter buffer (one launch per step).
// Reference: the drivers build torch.optim.Adamax(gnn.parameters(), lr) (scripts/main_gnn.py:160-167,
// scripts/main_generate.py:156-163); this is the same update rule (torch defaults beta1=0.9,
// beta2=0.999, eps=1e-8, no weight decay) applied to all parameters at once.  grad_scale folds the
// 1/world_size of the data-parallel gradient all-reduce into the same pass.
#include "common.cuh"

__global__ void adamax_bump_kernel(int* step) { *step += 1; }

__global__ void adamax_kernel(float* __restrict__ p, const float* __restrict__ g,
                              float* __restrict__ m, float* __restrict__ u, long long n, float lr,
                              float beta1, float beta2, float eps, float gscale,
                              const int* __restrict__ step) {
    // bias correction from the DEVICE step counter, so a captured CUDA graph stays correct on replay
    const float clr = lr / (1.f - powf(beta1, (float)step[0]));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float ui = fmaxf(beta2 * u[i], fabsf(gi) + eps);
        m[i] = mi;
        u[i] = ui;
        p[i] -= clr * mi / ui;
    }
}

extern "C" int hgnn_adamax_step(float* param, const float* grad, float* exp_avg, float* exp_inf,
                                long long n, float lr, float beta1, float beta2, float eps,
                                float grad_scale, int* step, hgnn_stream_t stream) {
    HGNN_REQUIRE(param && grad && exp_avg && exp_inf && n >= 0 && step, "bad argument");
    if (n == 0) return HGNN_OK;
    adamax_bump_kernel<<<1, 1, 0, to_stream(stream)>>>(step);
    int grid = persistent_grid(ceil_div(n, 256), 8);
    adamax_kernel<<<grid, 256, 0, to_stream(stream)>>>(param, grad, exp_avg, exp_inf, n, lr, beta1,
                                                       beta2, eps, grad_scale, step);
    return hgnn_check_launch("hgnn_adamax_step");
}
