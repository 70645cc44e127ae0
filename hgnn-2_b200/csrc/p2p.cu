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
// Buffer of one rank (hgnn_p2p_alloc): [flags: P2P_MAX_RANKS x u32 in a 128-byte header][slots: 2 parities x
// P2P_MAX_RANKS writers x cap4 floats] (cap4 = cap rounded up to a multiple of 4: 16-byte aligned slots).
//
// PUSH (default): in step s every rank WRITES its gradient into its own slot (s & 1, rank) of every PEER's buffer -
// posted stores over NVLink, nobody waits for a round trip - fences, and raises flag[rank] = s in every peer's header;
// then it waits for the flags the peers raised in ITS OWN header (local polls), reads their gradients from its own
// memory (ld.global.cg: the lines were written remotely, L1 must not serve them), sums them in rank order (every rank
// adds the same numbers in the same order, so the replicas stay bit-identical) and applies the update.  The pull
// variant of the first version (every rank publishes locally and reads the peers' buffers: a remote poll per flag and
// three rounds of scalar remote loads per thread; HGNN_B200_P2P_PULL=1) cost 13 us of pure latency at 2 ranks.
//
// Safety of the slot reuse (both variants): step s uses parity s & 1 and flag value s.  A rank can only finish step
// s + 1 after every peer has raised flag s + 1, which a peer does at the start of its step s + 1 kernel - i.e. after
// its step s kernel, with all its reads of step-s data, has completed.  Parity s & 1 is written again in step s + 2.
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

__device__ __forceinline__ long long p2p_cap4(long long cap) { return (cap + 3) & ~3ll; }
__device__ __forceinline__ float* p2p_slot(float* buf, int parity, int writer, long long cap4) {
    return buf + P2P_HEADER / 4 + ((size_t)parity * P2P_MAX_RANKS + writer) * cap4;
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
    const long long cap4 = p2p_cap4(cap);
    float* mine = p2p_slot(peers.buf[rank], s & 1, rank, cap4);
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
            const float* src = p2p_slot(peers.buf[r], s & 1, r, cap4);
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

__global__ void __launch_bounds__(P2P_THREADS)
p2p_push_allreduce_adamax_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ u, int n, float lr, float beta1, float beta2, float eps, float gscale,
                                 int* __restrict__ step, P2pPeers peers, int rank, int world, long long cap, int* fault, int fence_all) {
    __shared__ int s_step;
    const int tid = threadIdx.x;
    if (tid == 0) s_step = step[0] + 1;
    __syncthreads();
    const int s = s_step, parity = s & 1;
    const long long cap4 = p2p_cap4(cap);
    const int n4 = n >> 2;
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    // ---- 1. my gradient into my slot of every peer's buffer (posted remote stores)
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        float* dst = p2p_slot(peers.buf[q], parity, rank, cap4);
        if (vec) {
            for (int i = tid; i < n4; i += P2P_THREADS)
                reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(g)[i];
            for (int i = 4 * n4 + tid; i < n; i += P2P_THREADS) dst[i] = g[i];
        } else {
            for (int i = tid; i < n; i += P2P_THREADS) dst[i] = g[i];
        }
    }
    // The flags go up with st.release.sys by the threads below: a release is cumulative over everything that happens
    // before it, and the CTA barrier orders every thread's data stores before it - no system fence per thread (1 024
    // fences cost ~1.5 us of the exposed time).  HGNN_B200_P2P_FENCE=1 restores the per-thread fence.
    if (fence_all) __threadfence_system();
    __syncthreads();
    if (tid == 0) step[0] = s;
    // ---- 2. flag[rank] = s in every peer's header; 3. wait for flag[q] >= s in MY header (one thread per peer)
    if (tid < world && tid != rank) {
        unsigned* theirs = reinterpret_cast<unsigned*>(peers.buf[tid]) + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"((unsigned)s) : "memory");
        const unsigned* mine = reinterpret_cast<const unsigned*>(peers.buf[rank]) + tid;
        long long spins = 0;
        while ((int)(ld_acquire_sys(mine) - (unsigned)s) < 0) {
            if (++spins > (1ll << 24)) { atomicExch(fault, 1 + tid); break; }      // a dead peer must not hang the GPU
        }
    }
    __syncthreads();
    // ---- 4. sum in rank order from my own memory, 5. Adamax
    const float clr = lr / (1.f - powf(beta1, (float)s));
    float* base = peers.buf[rank];
    auto update = [&](int i, float sum) {
        const float gi = sum * gscale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float ui = fmaxf(beta2 * u[i], fabsf(gi) + eps);
        m[i] = mi;
        u[i] = ui;
        p[i] -= clr * mi / ui;
    };
    if (vec) {
        for (int i = tid; i < n4; i += P2P_THREADS) {
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < P2P_MAX_RANKS; ++r) {        // rank order; the loads are independent (unrolled: all in flight)
                if (r < world) {
                    const float4 v = (r == rank) ? reinterpret_cast<const float4*>(g)[i]
                                                 : __ldcg(reinterpret_cast<const float4*>(p2p_slot(base, parity, r, cap4)) + i);
                    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                }
            }
            update(4 * i, sum.x); update(4 * i + 1, sum.y); update(4 * i + 2, sum.z); update(4 * i + 3, sum.w);
        }
    }
    for (int i = (vec ? 4 * n4 : 0) + tid; i < n; i += P2P_THREADS) {
        float sum = 0.f;
        for (int r = 0; r < world; ++r) sum += (r == rank) ? g[i] : __ldcg(p2p_slot(base, parity, r, cap4) + i);
        update(i, sum);
    }
}

extern "C" long long hgnn_p2p_buffer_bytes(long long cap_floats) {
    return P2P_HEADER + 2ll * P2P_MAX_RANKS * ((cap_floats + 3) & ~3ll) * 4;
}
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
    static int pull = -1;
    if (pull < 0) { const char* e = getenv("HGNN_B200_P2P_PULL"); pull = (e && e[0] == '1') ? 1 : 0; }
    static int fence_all = -1;
    if (fence_all < 0) { const char* e = getenv("HGNN_B200_P2P_FENCE"); fence_all = (e && e[0] == '1') ? 1 : 0; }
    if (pull)
        p2p_allreduce_adamax_kernel<<<1, P2P_THREADS, 0, to_stream(stream)>>>(param, grad, exp_avg, exp_inf, n, lr, beta1,
                                                                              beta2, eps, grad_scale, step, peers, rank,
                                                                              world, cap_floats, fault);
    else
        p2p_push_allreduce_adamax_kernel<<<1, P2P_THREADS, 0, to_stream(stream)>>>(param, grad, exp_avg, exp_inf, n, lr,
                                                                                   beta1, beta2, eps, grad_scale, step,
                                                                                   peers, rank, world, cap_floats, fault,
                                                                                   fence_all);
    return hgnn_check_launch("hgnn_p2p_allreduce_adamax");
}
