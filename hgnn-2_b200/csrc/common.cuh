// common.cuh -- shared helpers for the hgnn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/hgnn_b200.h"

#define HGNN_SM_COUNT 148          // B200: 2 dies x 74 SMs
#define HGNN_MAX_GRID (HGNN_SM_COUNT * 8)

void hgnn_set_error(const char* fmt, ...);
int hgnn_check_launch(const char* what);
// Programmatic dependent launch for the next width-4 engine kernels issued by this thread (set by the
// step executor in program.cu around launches whose stream predecessor is another side kernel of the
// same chain: such a predecessor writes activations / accumulators only, never parameters or graph
// structure, which is all a PDL kernel touches before its griddepcontrol.wait).
void hgnn_eng_set_pdl(bool on);
// Launch recorder of the thread-per-row engine kernels (engine.cu: eng_launch; program.cu: replay of a pass as one
// graph launch).  While installed (thread-local), those kernels are not launched but appended here.
#include <vector>
struct hgnn_eng_slot_t {
    const void* func;
    int grid, block;
    unsigned smem;
    int pdl;
    std::vector<char> args;      // the kernel's single by-value parameter
};
struct hgnn_eng_recorder_t { std::vector<hgnn_eng_slot_t> slots; };
void hgnn_eng_set_recorder(hgnn_eng_recorder_t* r);

#define HGNN_REQUIRE(cond, msg)                      \
    do {                                             \
        if (!(cond)) {                               \
            hgnn_set_error("%s: %s", __func__, msg); \
            return HGNN_ERR_ARG;                     \
        }                                            \
    } while (0)

static inline cudaStream_t to_stream(hgnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid for a persistent / grid-stride kernel: enough CTAs to cover `items`, capped at a multiple
// of the SM count so every SM gets the same number of resident CTAs.
static inline int persistent_grid(long long items_per_cta_units, int ctas_per_sm) {
    long long cap = (long long)HGNN_SM_COUNT * ctas_per_sm;
    if (cap > HGNN_MAX_GRID) cap = HGNN_MAX_GRID;
    long long g = items_per_cta_units < 1 ? 1 : items_per_cta_units;
    return (int)(g < cap ? g : cap);
}

// Device-side copy of an operator list (passed by value as a kernel parameter).
struct OpList {
    int n;
    int kind[HGNN_MAX_OPS];
    const float* diag[HGNN_MAX_OPS];
    const int* rowptr[HGNN_MAX_OPS];
    const int* col[HGNN_MAX_OPS];
    const float* val[HGNN_MAX_OPS];
    const int* rng_rowptr[HGNN_MAX_OPS];
    const int* rng_id[HGNN_MAX_OPS];
    const float* rng_val[HGNN_MAX_OPS];
    const int* rng_lo[HGNN_MAX_OPS];
    const int* rng_hi[HGNN_MAX_OPS];
};

static inline int make_oplist(const hgnn_op_t* ops, int n_ops, OpList* out) {
    if (n_ops < 0 || n_ops > HGNN_MAX_OPS) return -1;
    out->n = n_ops;
    for (int i = 0; i < HGNN_MAX_OPS; ++i) {
        out->kind[i] = HGNN_OP_IDENT;
        out->diag[i] = nullptr;
        out->rowptr[i] = nullptr;
        out->col[i] = nullptr;
        out->val[i] = nullptr;
        out->rng_rowptr[i] = nullptr;
        out->rng_id[i] = nullptr;
        out->rng_val[i] = nullptr;
        out->rng_lo[i] = nullptr;
        out->rng_hi[i] = nullptr;
    }
    for (int i = 0; i < n_ops; ++i) {
        out->kind[i] = ops[i].kind;
        out->diag[i] = ops[i].diag;
        out->rowptr[i] = ops[i].rowptr;
        out->col[i] = ops[i].col;
        out->val[i] = ops[i].val;
        if (ops[i].kind == HGNN_OP_CSR && ops[i].rng_rowptr) {
            out->rng_rowptr[i] = ops[i].rng_rowptr;
            out->rng_id[i] = ops[i].rng_id;
            out->rng_val[i] = ops[i].rng_val;
            out->rng_lo[i] = ops[i].rng_lo;
            out->rng_hi[i] = ops[i].rng_hi;
        }
        if (ops[i].kind == HGNN_OP_DIAG && !ops[i].diag) return -1;
        if (ops[i].kind == HGNN_OP_CSR && (!ops[i].rowptr)) return -1;
        if (ops[i].kind < 0 || ops[i].kind > HGNN_OP_CSR) return -1;
    }
    return 0;
}

// ---- cross-CTA reduction ("last block done") -------------------------------------------------
// Workspace layout: [0, 256) bytes = ticket counter (self-resetting), then NB x width fp64
// accumulators (all zero on entry and on exit).  Every CTA adds its partial sums into bin
// blockIdx % NB with fire-and-forget fp64 atomics (binning keeps same-address contention at
// grid/NB); the CTA that takes the last ticket sums the bins in a fixed order, re-zeroes them and
// runs the finalizer in the same launch.
#define HGNN_WS_HEADER 256

// number of accumulator bins for a reduction of `width` values (host and device agree on this).  Every CTA of a launch
// ends with fire-and-forget fp64 reductions into its bin, and the launch is complete - its dependent released - only when
// the L2 atomic units have worked through all of them.  A launch of ~600 CTAs used to send 28 k reductions at the five
// cache lines of a single 80-value dW block: the dependent launch was released 1.8 us after the last CTA had exited instead
// of 0.7 us (profiles/logs/step_timeline_r3e.log vs _r3g).  So:
//   width <= 16 (the 2F batch-norm sums, dbias): bins x width = HGNN_WS_SMALL_DOUBLES = 32 doubles (width 8: 4 bins, one
//                load per lane of a consumer warp, hgnn_bins8_lane).  128 doubles (16 bins, four loads per lane) measured
//                SLOWER on C2: 0.629 vs 0.610 ms per step (profiles/logs/bench_r3i_*.log);
//   width <= 256 (the dW of a narrow side): 8 bins;
//   wider (dW of a wide side: one bin is already hundreds of cache lines): 1 bin.
#ifndef HGNN_WS_SMALL_DOUBLES
#define HGNN_WS_SMALL_DOUBLES 32
#endif
#define HGNN_WS_WIDE_BINS 8
__host__ __device__ inline int hgnn_ws_bins(int width) {
    if (width > 256) return 1;
    if (width > 16) return HGNN_WS_WIDE_BINS;
    int nb = 1;
    while (nb * 2 * width <= HGNN_WS_SMALL_DOUBLES) nb *= 2;
    return nb;
}

__device__ __forceinline__ void accum_add(double* accum, int width, int nb, int idx, double v) {
    atomicAdd(accum + (size_t)(blockIdx.x & (nb - 1)) * width + idx, v);
}

// One warp reads a width-8 accumulator block (hgnn_ws_bins(8) bins x 8 doubles): lane l sums the entries l, l + 32, ... -
// independent loads - so that lanes with equal (l & 7) hold parts of column l & 7; fold them with xor 8 and xor 16.
__device__ __forceinline__ double hgnn_bins8_lane(const double* __restrict__ acc) {
    constexpr int N = HGNN_WS_SMALL_DOUBLES / 32;
    const int lane = threadIdx.x & 31;
    double part[N];
#pragma unroll
    for (int j = 0; j < N; ++j) part[j] = __ldcg(acc + lane + 32 * j);
    double v = part[0];
#pragma unroll
    for (int j = 1; j < N; ++j) v += part[j];
    return v;
}

// last CTA only: total of column idx over the bins (fixed order), re-zeroing as it goes
__device__ __forceinline__ double accum_take(double* accum, int width, int nb, int idx) {
    double t0 = 0.0, t1 = 0.0;
    for (int b = 0; b < nb; b += 2) {
        double* p0 = accum + (size_t)b * width + idx;
        t0 += __ldcg(p0);
        *p0 = 0.0;
        if (b + 1 < nb) {
            double* p1 = p0 + width;
            t1 += __ldcg(p1);
            *p1 = 0.0;
        }
    }
    return t0 + t1;
}

__device__ __forceinline__ bool last_block_ticket(unsigned int* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Last-CTA reduction of per-CTA partial vectors: out[c] = sum_p partial[p*width + c], p in CTA
// order.  All threads take part (column c is split into blockDim/width row chunks that are then
// combined in a fixed order), so the tail costs ~nparts/chunks dependent L2 loads per thread
// instead of nparts.  `scratch` = blockDim.x elements of shared memory; `out` may be shared or
// global.  Ends with a __syncthreads().
template <typename T>
__device__ __forceinline__ void lastblock_reduce(const T* __restrict__ partial, int nparts, int width,
                                                 T* out, T* scratch) {
    const int nthr = blockDim.x;
    for (int c0 = 0; c0 < width; c0 += nthr) {
        const int w = min(nthr, width - c0);
        const int nch = nthr / w;
        const int c = threadIdx.x % w, ch = threadIdx.x / w;
        T acc = (T)0;
        if (ch < nch) {
            int p = ch;
            T a0 = (T)0, a1 = (T)0, a2 = (T)0, a3 = (T)0;   // four loads in flight
            for (; p + 3 * nch < nparts; p += 4 * nch) {
                a0 += __ldcg(partial + (size_t)p * width + c0 + c);
                a1 += __ldcg(partial + (size_t)(p + nch) * width + c0 + c);
                a2 += __ldcg(partial + (size_t)(p + 2 * nch) * width + c0 + c);
                a3 += __ldcg(partial + (size_t)(p + 3 * nch) * width + c0 + c);
            }
            for (; p < nparts; p += nch) a0 += __ldcg(partial + (size_t)p * width + c0 + c);
            acc = (a0 + a1) + (a2 + a3);
        }
        scratch[threadIdx.x] = acc;
        __syncthreads();
        if ((int)threadIdx.x < w) {
            T t = (T)0;
            for (int k = 0; k < nch; ++k) t += scratch[k * w + threadIdx.x];
            out[c0 + threadIdx.x] = t;
        }
        __syncthreads();
    }
}

// Sum `v` over all threads of the CTA that share (threadIdx.x % M); M must divide 32.  Threads
// tid < M return the total of column tid; `sm` = (blockDim/32)*32 doubles.  Two __syncthreads().
__device__ __forceinline__ double cta_reduce_mod(double v, int M, double* sm) {
    for (int off = 16; off >= M; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane < M) sm[warp * 32 + lane] = v;
    __syncthreads();
    double t = 0.0;
    if ((int)threadIdx.x < M)
        for (int w = 0; w < nw; ++w) t += sm[w * 32 + threadIdx.x];
    return t;
}

// grid that gives every CTA the same number of tiles (no ragged second wave)
static inline int balanced_grid(int ntiles, int max_ctas) {
    if (ntiles <= max_ctas) return ntiles < 1 ? 1 : ntiles;
    int per = (ntiles + max_ctas - 1) / max_ctas;
    return (ntiles + per - 1) / per;
}
