// gmul.cu -- stand-alone multi-operator aggregation ("gmul"), forward and backward.
//
// Reference: graph_oper.forward / graph_op (models/layers/layers_mnb.py:395-411,
// functions/utils.py:24-52) and P_multi.forward / Pmul (:418-434, :55-81): a Python loop of dense
// torch.mm over operators that are >= 99 % zeros.  Here ONE pass over the packed feature rows
// produces all operator blocks of a row: a group of LPR lanes owns a row, lane l owns features
// l, l+LPR, ...; the row's CSR segment is read once per operator (same address across the group =
// one broadcast transaction) and each neighbour row is gathered with consecutive lanes on
// consecutive floats (coalesced).  HBM-bound; no tensor cores.
#include "common.cuh"

template <int LPR, bool BWD>
__global__ void __launch_bounds__(256)
gmul_kernel(OpList ops, int R, int F, const float* __restrict__ X, float* __restrict__ Y) {
    constexpr int ROWS_PER_CTA = 256 / LPR;
    const int l = threadIdx.x % LPR;
    const int rl = threadIdx.x / LPR;
    const int K = ops.n;
    // FWD: X is (R_in, F); Y is (R, K*F), block t <- op t applied to X.
    // BWD: X is G (R_in, K*F); Y is gX (R, F) = sum_t op_t applied to block t of G.
    const int ldx = BWD ? K * F : F;
    const int ldy = BWD ? F : K * F;
    for (int r = blockIdx.x * ROWS_PER_CTA + rl; r < R; r += gridDim.x * ROWS_PER_CTA) {
        for (int f = l; f < F; f += LPR) {
            float total = 0.f;
#pragma unroll 1
            for (int t = 0; t < K; ++t) {
                const int xoff = BWD ? t * F + f : f;
                float acc;
                const int kind = ops.kind[t];
                if (kind == HGNN_OP_IDENT) {
                    acc = X[(size_t)r * ldx + xoff];
                } else if (kind == HGNN_OP_DIAG) {
                    acc = ops.diag[t][r] * X[(size_t)r * ldx + xoff];
                } else {
                    const int* __restrict__ col = ops.col[t];
                    const float* __restrict__ val = ops.val[t];
                    const int k0 = ops.rowptr[t][r], k1 = ops.rowptr[t][r + 1];
                    acc = 0.f;
                    int k = k0;
                    for (; k + 1 < k1; k += 2) {   // two independent gathers in flight
                        float a0 = val[k] * X[(size_t)col[k] * ldx + xoff];
                        float a1 = val[k + 1] * X[(size_t)col[k + 1] * ldx + xoff];
                        acc += a0;
                        acc += a1;
                    }
                    if (k < k1) acc += val[k] * X[(size_t)col[k] * ldx + xoff];
                }
                if (BWD) total += acc;
                else Y[(size_t)r * ldy + t * F + f] = acc;
            }
            if (BWD) Y[(size_t)r * ldy + f] = total;
        }
    }
}

template <bool BWD>
static int launch_gmul(const hgnn_op_t* ops, int n_ops, int R, int F, const float* X, float* Y,
                       hgnn_stream_t stream) {
    OpList ol;
    HGNN_REQUIRE(ops && n_ops >= 1 && make_oplist(ops, n_ops, &ol) == 0, "bad operator list");
    HGNN_REQUIRE(R >= 0 && F >= 1 && X && Y, "bad argument");
    if (R == 0) return HGNN_OK;
    cudaStream_t s = to_stream(stream);
#define LAUNCH(LPR)                                                                         \
    {                                                                                       \
        int grid = persistent_grid(ceil_div(R, 256 / LPR), 8);                              \
        gmul_kernel<LPR, BWD><<<grid, 256, 0, s>>>(ol, R, F, X, Y);                         \
    }
    if (F <= 1) LAUNCH(1)
    else if (F <= 2) LAUNCH(2)
    else if (F <= 4) LAUNCH(4)
    else if (F <= 8) LAUNCH(8)
    else if (F <= 16) LAUNCH(16)
    else LAUNCH(32)
#undef LAUNCH
    return hgnn_check_launch(BWD ? "hgnn_gmul_bwd" : "hgnn_gmul_fwd");
}

extern "C" int hgnn_gmul_fwd(const hgnn_op_t* ops, int n_ops, int R, int F, const float* X,
                             float* Y, hgnn_stream_t stream) {
    return launch_gmul<false>(ops, n_ops, R, F, X, Y, stream);
}

extern "C" int hgnn_gmul_bwd(const hgnn_op_t* opsT, int n_ops, int R, int F, const float* G,
                             float* gX, hgnn_stream_t stream) {
    return launch_gmul<true>(opsT, n_ops, R, F, G, gX, stream);
}
