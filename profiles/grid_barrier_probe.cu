// grid_barrier_probe.cu -- what does one grid-wide barrier cost on B200?  Input for the decision whether the h = 2
// step (76 dependent launches of 7-18 us against a 4-7 us launch floor, profiles/README.md) should become one
// persistent kernel per pass with a barrier between layer sides.  Stand-alone:
//     nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gb_probe profiles/grid_barrier_probe.cu
//     timeout 30 /tmp/gb_probe
// Prints us per barrier for 1..8 CTAs per SM (128 threads each), for (a) a sense-reversing barrier on one global
// counter (atomicAdd + ld.acquire spin by one thread per CTA), (b) the same with a 32-byte "payload" reduction
// (fp64 atomics into 4 accumulators before the barrier, read after it: what a batch-norm statistic needs), and
// (c) back-to-back empty kernel launches inside a CUDA graph for comparison.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int* gen, unsigned int nblocks, unsigned int& local_gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++local_gen;
        __threadfence();
        if (atomicAdd(counter, 1u) == nblocks - 1) {
            *counter = 0;
            __threadfence();
            atomicExch(gen, local_gen);            // release everybody
        } else {
            unsigned int g;
            do {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(gen) : "memory");
            } while (g != local_gen);
        }
    }
    __syncthreads();
}

__global__ void barrier_kernel(unsigned int* counter, unsigned int* gen, int iters, double* acc, int payload, double* sink) {
    unsigned int local_gen = 0;
    double s = 0.0;
    for (int i = 0; i < iters; ++i) {
        if (payload && threadIdx.x < 4) atomicAdd(acc + (i & 1) * 4 + threadIdx.x, 1.0);
        grid_barrier(counter, gen, gridDim.x, local_gen);
        if (payload && threadIdx.x < 4) s += __ldcg(acc + (i & 1) * 4 + threadIdx.x);
    }
    if (payload && threadIdx.x < 4 && blockIdx.x == 0) sink[threadIdx.x] = s;
}

__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 1234567) p[0] = 1; }

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned int* ctr; double *acc, *sink;
    cudaMalloc(&ctr, 256); cudaMalloc(&acc, 64); cudaMalloc(&sink, 32);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int payload = 0; payload < 2; ++payload)
        for (int per = 1; per <= 8; per *= 2) {
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, barrier_kernel, 128, 0);
            if (occ < per) break;
            cudaMemset(ctr, 0, 256); cudaMemset(acc, 0, 64);
            void* args[] = {(void*)&ctr, nullptr, (void*)&iters, (void*)&acc, (void*)&payload, (void*)&sink};
            unsigned int* gen = ctr + 32;
            args[1] = (void*)&gen;
            // cooperative launch only to guarantee co-residency of the whole grid
            cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(sms * per), dim3(128), args, 0, 0);   // warm-up
            cudaDeviceSynchronize();
            cudaMemset(ctr, 0, 256);
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void*)barrier_kernel, dim3(sms * per), dim3(128), args, 0, 0);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            printf("%s barrier, %d CTAs/SM (%d CTAs): %.2f us per barrier  [%s]\n", payload ? "payload" : "plain  ",
                   per, sms * per, ms * 1e3 / iters, cudaGetErrorString(e));
        }
    // back-to-back dependent launches inside a graph
    cudaStream_t st; cudaStreamCreate(&st);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 500; ++i) empty_kernel<<<sms * 4, 128, 0, st>>>(nullptr);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("500 dependent empty launches of %d CTAs in a graph: %.2f us per launch\n", sms * 4, ms * 1e3 / 500);
    return 0;
}
