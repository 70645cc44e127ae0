// graph_update_probe.cu -- host cost of re-parametrising a captured kernel chain against issuing the launches.
// Decides whether the per-step launch sequence of the step executor (~40 dependent side kernels per pass, different
// sizes and pointers every batch) should be replayed as ONE graph launch after cudaGraphExecKernelNodeSetParams on
// every node, instead of ~40 cudaLaunchKernelEx calls (4.6 us each on the box, profiles/logs/e2e_phases_r2k.log).
//     nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gu_probe profiles/graph_update_probe.cu && /tmp/gu_probe
#include <chrono>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>

struct Args { const float* a; float* b; int n; int pad[93]; };     // ~400 bytes, like Bwd4Args
__global__ void k(const Args p) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < p.n) p.b[i] = p.a[i] + 1.f; }

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
    const int N = 40, iters = 300;
    float *a, *b;
    cudaMalloc(&a, 1 << 20); cudaMalloc(&b, 1 << 20);
    cudaStream_t s; cudaStreamCreate(&s);
    Args p{a, b, 1000};
    // (1) direct launches with the programmatic-serialization attribute
    auto launch = [&](int n) {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(8); cfg.blockDim = dim3(128); cfg.stream = s;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = 1;
        p.n = n;
        cudaLaunchKernelEx(&cfg, k, p);
    };
    for (int i = 0; i < N; ++i) launch(1000);
    cudaStreamSynchronize(s);
    double t0 = now();
    for (int it = 0; it < iters; ++it) { for (int i = 0; i < N; ++i) launch(1000 + it); cudaStreamSynchronize(s); }
    double t_direct = (now() - t0) / iters;
    // (2) captured chain, every node re-parametrised, one graph launch
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < N; ++i) launch(1000);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    size_t nn = 0; cudaGraphGetNodes(g, nullptr, &nn);
    std::vector<cudaGraphNode_t> nodes(nn); cudaGraphGetNodes(g, nodes.data(), &nn);
    std::vector<cudaGraphNode_t> kn;
    for (auto n : nodes) { cudaGraphNodeType t; cudaGraphNodeGetType(n, &t); if (t == cudaGraphNodeTypeKernel) kn.push_back(n); }
    printf("graph nodes %zu, kernel nodes %zu\n", nn, kn.size());
    cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    t0 = now();
    cudaError_t e = cudaSuccess;
    for (int it = 0; it < iters; ++it) {
        for (size_t i = 0; i < kn.size(); ++i) {
            Args q{a, b, 1000 + it};
            void* kargs[] = {&q};
            cudaKernelNodeParams kp = {};
            kp.func = (void*)k; kp.gridDim = dim3(8 + (it & 1)); kp.blockDim = dim3(128); kp.sharedMemBytes = 0; kp.kernelParams = kargs;
            cudaError_t r = cudaGraphExecKernelNodeSetParams(ge, kn[i], &kp);
            if (r != cudaSuccess) e = r;
        }
        cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    }
    double t_update = (now() - t0) / iters;
    // (3) graph launch alone
    t0 = now();
    for (int it = 0; it < iters; ++it) { cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); }
    double t_replay = (now() - t0) / iters;
    printf("%d launches incl. sync: direct %.1f us | set-params on every node + graph launch %.1f us | graph launch only %.1f us   [%s]\n",
           N, t_direct * 1e6, t_update * 1e6, t_replay * 1e6, cudaGetErrorString(e));
    // host time only (no sync in the loop)
    t0 = now();
    for (int it = 0; it < iters; ++it) for (int i = 0; i < N; ++i) launch(1000 + it);
    double h_direct = (now() - t0) / iters; cudaStreamSynchronize(s);
    t0 = now();
    for (int it = 0; it < iters; ++it) {
        for (size_t i = 0; i < kn.size(); ++i) {
            Args q{a, b, 1000 + it}; void* kargs[] = {&q};
            cudaKernelNodeParams kp = {}; kp.func = (void*)k; kp.gridDim = dim3(8); kp.blockDim = dim3(128); kp.kernelParams = kargs;
            cudaGraphExecKernelNodeSetParams(ge, kn[i], &kp);
        }
        cudaGraphLaunch(ge, s);
    }
    double h_update = (now() - t0) / iters; cudaStreamSynchronize(s);
    printf("host issue time only: direct %.1f us | set-params + graph launch %.1f us\n", h_direct * 1e6, h_update * 1e6);
    return 0;
}
