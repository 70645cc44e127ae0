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
    }
    for (int i = 0; i < n_ops; ++i) {
        out->kind[i] = ops[i].kind;
        out->diag[i] = ops[i].diag;
        out->rowptr[i] = ops[i].rowptr;
        out->col[i] = ops[i].col;
        out->val[i] = ops[i].val;
        if (ops[i].kind == HGNN_OP_DIAG && !ops[i].diag) return -1;
        if (ops[i].kind == HGNN_OP_CSR && (!ops[i].rowptr)) return -1;
        if (ops[i].kind < 0 || ops[i].kind > HGNN_OP_CSR) return -1;
    }
    return 0;
}

// ---- cross-CTA deterministic reduction ("last block done") ---------------------------------
// Each CTA writes its partial vector to ws[cta][width]; the CTA that takes the last ticket
// reduces all partials in CTA order (fixed order => bit-reproducible) and runs the finalizer.
// Workspace layout: [0, 256) bytes = ticket counter (self-resetting), then partials.
#define HGNN_WS_HEADER 256

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
