// p2p.cu -- gradient all-reduce FUSED with the Adamax update, over NVLink peer memory.
//
// The only exchange of the data-parallel path is the sum of one flat fp32 gradient per step (SURVEY.md 8e:
// ~3 k floats for LGNN L = 20, h = 2 - latency-bound).  Through NCCL that collective sits fully exposed between
// the last backward kernel and the optimizer: +19 us at 2 GPUs, +75 us at 8 (SCALE_r01: 0.923 efficiency).
// Here every rank publishes its gradient in a buffer its peers can read (CUDA IPC, NVSwitch: every GPU reaches
// every peer at full bandwidth), raises a flag, and ONE kernel per rank waits for the peers' flags, reads their
// gradients straight over NVLink, sums them in rank order (every rank adds the same numbers in the same order, so
// the replicas stay bit-identical) and applies torch.optim.Adamax's rule (scripts/main_gnn.py:160-167) to the
// flat parameter buffer - no separate collective, no extra pass over the gradient.
//
// Buffer of one rank (hgnn_p2p_alloc): [flag: u32, 128-byte padded][slot 0: cap floats][slot 1: cap floats].
// Step s uses slot s & 1 and flag value s: a rank can only start step s + 1 after it has read every peer's step-s
// data, and a peer overwrites slot s & 1 in step s + 2 at the earliest - after this rank's step s + 1 flag, which
// it raises after its step-s reads.  Peer data is read with ld.volatile (peer lines must not be served from L1).
#include "common.cuh"

#define P2P_THREADS 1024
#define P2P_MAX_RANKS 16
#define P2P_HEADER 128

struct P2pPeers { float* buf[P2P_MAX_RANKS]; };

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_adamax_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ u, int n, float lr, float beta1, float beta2, float eps, float gscale,
                            int* __restrict__ step, P2pPeers peers, int rank, int world, long long cap, int* fault) {
    __shared__ int s_step;
    const int tid = threadIdx.x;
    if (tid == 0) s_step = step[0] + 1;
    __syncthreads();
    const int s = s_step;
    float* mine = peers.buf[rank] + P2P_HEADER / 4 + (size_t)(s & 1) * cap;
    for (int i = tid; i < n; i += P2P_THREADS) mine[i] = g[i];
    __syncthreads();
    if (tid == 0) {
        step[0] = s;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<unsigned*>(peers.buf[rank])), "r"((unsigned)s) : "memory");
    }
    // one thread per peer waits for that peer's flag (bounded: a dead peer must not hang the GPU)
    if (tid < world && tid != rank) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(peers.buf[tid]);
        long long spins = 0;
        while ((int)(ld_acquire_sys(flag) - (unsigned)s) < 0) {
            if (++spins > (1ll << 23)) { atomicExch(fault, 1 + tid); break; }
        }
    }
    __syncthreads();
    const float clr = lr / (1.f - powf(beta1, (float)s));
    for (int i = tid; i < n; i += P2P_THREADS) {
        float sum = 0.f;
        for (int r = 0; r < world; ++r) {
            const float* src = peers.buf[r] + P2P_HEADER / 4 + (size_t)(s & 1) * cap;
            sum += (r == rank) ? g[i] : ld_volatile_f32(src + i);
        }
        const float gi = sum * gscale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float ui = fmaxf(beta2 * u[i], fabsf(gi) + eps);
        m[i] = mi;
        u[i] = ui;
        p[i] -= clr * mi / ui;
    }
}

extern "C" long long hgnn_p2p_buffer_bytes(long long cap_floats) { return P2P_HEADER + 2 * cap_floats * 4; }
extern "C" int hgnn_p2p_max_floats(void) { return P2P_THREADS * 16; }

extern "C" int hgnn_p2p_alloc(long long cap_floats, void** dev_ptr, void* handle64) {
    HGNN_REQUIRE(cap_floats > 0 && dev_ptr && handle64, "bad argument");
    void* ptr = nullptr;
    const size_t bytes = (size_t)hgnn_p2p_buffer_bytes(cap_floats);
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e == cudaSuccess) e = cudaMemset(ptr, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        hgnn_set_error("hgnn_p2p_alloc: %s", cudaGetErrorString(e));
        if (ptr) cudaFree(ptr);
        return HGNN_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *dev_ptr = ptr;
    return HGNN_OK;
}

extern "C" int hgnn_p2p_open(const void* handle64, void** dev_ptr) {
    HGNN_REQUIRE(handle64 && dev_ptr, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        hgnn_set_error("hgnn_p2p_open: %s", cudaGetErrorString(e));
        return HGNN_ERR_CUDA;
    }
    return HGNN_OK;
}

extern "C" int hgnn_p2p_close(void* dev_ptr) { return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? HGNN_OK : HGNN_ERR_CUDA; }
extern "C" int hgnn_p2p_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? HGNN_OK : HGNN_ERR_CUDA; }

extern "C" int hgnn_p2p_allreduce_adamax(float* param, const float* grad, float* exp_avg, float* exp_inf, int n,
                                         float lr, float beta1, float beta2, float eps, float grad_scale, int* step,
                                         void* const* peer_bufs, int rank, int world, long long cap_floats, int* fault,
                                         hgnn_stream_t stream) {
    HGNN_REQUIRE(param && grad && exp_avg && exp_inf && step && peer_bufs && fault, "null argument");
    HGNN_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world size");
    HGNN_REQUIRE(n >= 0 && n <= cap_floats && n <= hgnn_p2p_max_floats(), "gradient larger than the peer buffer");
    if (n == 0) return HGNN_OK;
    P2pPeers peers;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) peers.buf[r] = r < world ? static_cast<float*>(peer_bufs[r]) : nullptr;
    for (int r = 0; r < world; ++r) HGNN_REQUIRE(peers.buf[r], "null peer buffer");
    p2p_allreduce_adamax_kernel<<<1, P2P_THREADS, 0, to_stream(stream)>>>(param, grad, exp_avg, exp_inf, n, lr, beta1,
                                                                          beta2, eps, grad_scale, step, peers, rank,
                                                                          world, cap_floats, fault);
    return hgnn_check_launch("hgnn_p2p_allreduce_adamax");
}
