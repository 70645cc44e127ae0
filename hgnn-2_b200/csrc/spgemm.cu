// spgemm.cu -- C = A * B on block-diagonal CSR, for the power operators A^(2^j) / AL^(2^j).
//
// Reference: functions/operators.py:26-29 and :78-81 square DENSE matrices with torch.matmul
// (O(N^3) / O(M^3)); here rows are short (nnz/row ~ d^2), so each warp owns one output row:
//   1. count the partial products of every row (host scans the counts),
//   2. expand them into scratch and flag the first occurrence of every output column,
//   3. merge equal columns in a fixed order and emit the row sorted by column.
// Entries are small integers / dyadic rationals, so the sums are exact in fp32
// and the result is bit-identical to the dense product.  clip=1 binarises (north_star's
// "SpGEMM-then-clip"); the reference itself does not clip, so the host default is clip=0.
#include "common.cuh"

__global__ void spgemm_count_kernel(int R, const int* __restrict__ a_rowptr,
                                    const int* __restrict__ a_col,
                                    const int* __restrict__ b_rowptr, int* __restrict__ prodcnt) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        int n = 0;
        for (int k = a_rowptr[r]; k < a_rowptr[r + 1]; ++k) {
            int c = a_col[k];
            n += b_rowptr[c + 1] - b_rowptr[c];
        }
        prodcnt[r] = n;
    }
}

extern "C" int hgnn_spgemm_count_products(int R, const int* a_rowptr, const int* a_col,
                                          const int* b_rowptr, int* prodcnt, hgnn_stream_t stream) {
    HGNN_REQUIRE(R >= 0 && a_rowptr && b_rowptr && prodcnt, "bad argument");
    if (R == 0) return HGNN_OK;
    spgemm_count_kernel<<<min(ceil_div(R, 256), HGNN_MAX_GRID), 256, 0, to_stream(stream)>>>(
        R, a_rowptr, a_col, b_rowptr, prodcnt);
    return hgnn_check_launch("hgnn_spgemm_count_products");
}

__global__ void __launch_bounds__(256)
spgemm_expand_kernel(int R, const int* __restrict__ a_rowptr, const int* __restrict__ a_col,
                     const float* __restrict__ a_val, const int* __restrict__ b_rowptr,
                     const int* __restrict__ b_col, const float* __restrict__ b_val,
                     const int* __restrict__ prodptr, int* pcol, float* pval, int* pflag,
                     int* __restrict__ rowcnt) {
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    for (int r = blockIdx.x * wpc + (threadIdx.x >> 5); r < R; r += gridDim.x * wpc) {
        const int base = prodptr[r], L = prodptr[r + 1] - base;
        int o = base;
        for (int k = a_rowptr[r]; k < a_rowptr[r + 1]; ++k) {
            const int c = a_col[k];
            const float av = a_val[k];
            const int b0 = b_rowptr[c], len = b_rowptr[c + 1] - b0;
            for (int j = lane; j < len; j += 32) {
                pcol[o + j] = b_col[b0 + j];
                pval[o + j] = av * b_val[b0 + j];
            }
            o += len;
        }
        __syncwarp();
        int cnt = 0;
        for (int p = lane; p < L; p += 32) {
            const int cp = pcol[base + p];
            int first = 1;
            for (int q = 0; q < p; ++q)
                if (pcol[base + q] == cp) { first = 0; break; }
            pflag[base + p] = first;
            cnt += first;
        }
        cnt = (int)__reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) rowcnt[r] = cnt;
    }
}

extern "C" int hgnn_spgemm_expand(int R, const int* a_rowptr, const int* a_col, const float* a_val,
                                  const int* b_rowptr, const int* b_col, const float* b_val,
                                  const int* prodptr, int* pcol, float* pval, int* pflag,
                                  int* rowcnt, hgnn_stream_t stream) {
    HGNN_REQUIRE(R >= 0 && a_rowptr && b_rowptr && prodptr && rowcnt, "bad argument");
    if (R == 0) return HGNN_OK;
    spgemm_expand_kernel<<<min(ceil_div(R, 8), HGNN_MAX_GRID), 256, 0, to_stream(stream)>>>(
        R, a_rowptr, a_col, a_val, b_rowptr, b_col, b_val, prodptr, pcol, pval, pflag, rowcnt);
    return hgnn_check_launch("hgnn_spgemm_expand");
}

__global__ void __launch_bounds__(256)
spgemm_fill_kernel(int R, const int* __restrict__ prodptr, const int* __restrict__ pcol,
                   const float* __restrict__ pval, const int* __restrict__ pflag,
                   const int* __restrict__ c_rowptr, int* __restrict__ c_col,
                   float* __restrict__ c_val, int clip) {
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    for (int r = blockIdx.x * wpc + (threadIdx.x >> 5); r < R; r += gridDim.x * wpc) {
        const int base = prodptr[r], L = prodptr[r + 1] - base;
        const int out0 = c_rowptr[r];
        for (int p = lane; p < L; p += 32) {
            if (!pflag[base + p]) continue;
            const int cp = pcol[base + p];
            int pos = 0;
            float sum = 0.f;
            for (int q = 0; q < L; ++q) {       // fixed order: deterministic (and exact anyway)
                const int cq = pcol[base + q];
                if (cq == cp) sum += pval[base + q];
                else if (cq < cp && pflag[base + q]) ++pos;
            }
            if (clip) sum = (sum != 0.f) ? 1.f : 0.f;
            c_col[out0 + pos] = cp;
            c_val[out0 + pos] = sum;
        }
    }
}

extern "C" int hgnn_spgemm_fill(int R, const int* prodptr, const int* pcol, const float* pval,
                                const int* pflag, const int* c_rowptr, int* c_col, float* c_val,
                                int clip, hgnn_stream_t stream) {
    HGNN_REQUIRE(R >= 0 && prodptr && c_rowptr, "bad argument");
    if (R == 0) return HGNN_OK;
    spgemm_fill_kernel<<<min(ceil_div(R, 8), HGNN_MAX_GRID), 256, 0, to_stream(stream)>>>(
        R, prodptr, pcol, pval, pflag, c_rowptr, c_col, c_val, clip);
    return hgnn_check_launch("hgnn_spgemm_fill");
}
