// ccn.cu -- CCN covariant contraction kernels (second order and the 1-D variant).
//
// Reference: CompnetUtils.update_F (functions/utils_ccn.py:281-300): for every vertex i, promote
// each neighbour's state with chi F chi^T (:225-239, two dense matmuls per neighbour), stack to
// T (n,n,n,C), form the rank-6 product T (x) adj (:57-63, n^5*C floats) and run 18 permute / mask /
// triple-sum contractions (functions/contraction.py:106-121) before Linear(18C -> H) + ReLU.
//
// Here one CTA owns one vertex.  chi is a partial permutation, so promotion is an index map
// m[a][b] = position of r_b inside the receptive field of r_a (binary search, never a matrix);
// adj_i = chis[i][i] is the identity (:293), so all 18 contractions collapse to eight partial
// sums of T which are accumulated from ONE pass over the neighbours'
// tiles into shared memory:
//     Sc[a,b] = sum_c T[a,b,c]     Sa[b,c] = sum_a T[a,b,c]     D1[a,d] = T[a,d,d]
//     D2[b,d] = T[d,b,d]           Sbc[a], Sac[b], Sall, Sd = sum_a T[a,a,a]
// and the Linear + ReLU epilogue reads them straight from shared memory (blocks 7..15 are the
// same tensor nine times: their weights are summed instead).  HBM traffic is the gathered tiles
// (sum_j d_j^2 C floats) plus the d_i^2 H output - O(n^3 C) work instead of O(18 * 3 * n^5 C).
#include "common.cuh"

int hgnn_grid_cap(int width);

#define CCN_THREADS 128
#define CCN_MAX_SMEM (200 * 1024)

__device__ __forceinline__ int find_pos(const int* __restrict__ nbr, int lo, int hi, int key) {
    // position of `key` in the sorted slice nbr[lo:hi), or -1
    int l = lo, h = hi;
    while (l < h) {
        int mid = (l + h) >> 1;
        int v = __ldg(nbr + mid);
        if (v < key) l = mid + 1; else h = mid;
    }
    return (l < hi && __ldg(nbr + l) == key) ? (l - lo) : -1;
}

struct CcnSmem {
    int* r;      // [nmax]
    int* m;      // [nmax*nmax]
    float* Sc;   // [n*n*C]
    float* Sa;
    float* D1;
    float* D2;
    float* Sbc;  // [n*C]
    float* Sac;  // [n*C]
    float* Sall; // [C]
    float* Sd;   // [C]
    float* extra;
};

__host__ __device__ inline size_t ccn2_smem_floats(int nmax, int C) {
    return (size_t)nmax + (size_t)nmax * nmax + 4 * (size_t)nmax * nmax * C + 2 * (size_t)nmax * C + 2 * C;
}

__device__ __forceinline__ CcnSmem carve(float* base, int nmax, int C) {
    CcnSmem s;
    s.r = reinterpret_cast<int*>(base);
    s.m = s.r + nmax;
    s.Sc = reinterpret_cast<float*>(s.m + nmax * nmax);
    s.Sa = s.Sc + nmax * nmax * C;
    s.D1 = s.Sa + nmax * nmax * C;
    s.D2 = s.D1 + nmax * nmax * C;
    s.Sbc = s.D2 + nmax * nmax * C;
    s.Sac = s.Sbc + nmax * C;
    s.Sall = s.Sac + nmax * C;
    s.Sd = s.Sall + C;
    s.extra = s.Sd + C;
    return s;
}

// Partial sums of T for vertex i into shared memory (ends with a __syncthreads()).
__device__ void ccn2_partial_sums(const CcnSmem& s, int i, const int* __restrict__ nbr_ptr,
                                  const int* __restrict__ nbr, const long long* __restrict__ f_off,
                                  const float* __restrict__ Fprev, int C, int n) {
    const int tid = threadIdx.x;
    const int base = nbr_ptr[i];
    for (int k = tid; k < n; k += CCN_THREADS) s.r[k] = nbr[base + k];
    __syncthreads();
    for (int k = tid; k < n * n; k += CCN_THREADS) {
        const int a = k / n, b = k - a * n;
        const int j = s.r[a];
        s.m[k] = find_pos(nbr, nbr_ptr[j], nbr_ptr[j + 1], s.r[b]);
    }
    __syncthreads();
    for (int k = tid; k < n * n * C; k += CCN_THREADS) {
        const int ch = k % C, ab = k / C;
        const int a = ab / n, b = ab - a * n;
        // Sc[a,b], D1[a,b] = T[a,b,b]
        {
            const int j = s.r[a];
            const int dj = nbr_ptr[j + 1] - nbr_ptr[j];
            const float* Fj = Fprev + f_off[j] * C;
            const int pb = s.m[a * n + b];
            float sc = 0.f, d1 = 0.f;
            if (pb >= 0) {
                for (int c = 0; c < n; ++c) {
                    const int pc = s.m[a * n + c];
                    if (pc >= 0) sc += __ldg(Fj + ((size_t)pb * dj + pc) * C + ch);
                }
                d1 = __ldg(Fj + ((size_t)pb * dj + pb) * C + ch);
            }
            s.Sc[k] = sc;
            s.D1[k] = d1;
        }
        // Sa[b', c'] with (b', c') = (a, b) of this slot;  D2[b', d] = T[d, b', d]
        {
            const int bb = a, cc = b;
            float sa = 0.f;
            for (int aa = 0; aa < n; ++aa) {
                const int pb = s.m[aa * n + bb], pc = s.m[aa * n + cc];
                if (pb >= 0 && pc >= 0) {
                    const int j = s.r[aa];
                    const int dj = nbr_ptr[j + 1] - nbr_ptr[j];
                    sa += __ldg(Fprev + f_off[j] * C + ((size_t)pb * dj + pc) * C + ch);
                }
            }
            s.Sa[k] = sa;
            const int pb = s.m[cc * n + bb], pd = s.m[cc * n + cc];
            float d2 = 0.f;
            if (pb >= 0 && pd >= 0) {
                const int j = s.r[cc];
                const int dj = nbr_ptr[j + 1] - nbr_ptr[j];
                d2 = __ldg(Fprev + f_off[j] * C + ((size_t)pb * dj + pd) * C + ch);
            }
            s.D2[k] = d2;
        }
    }
    __syncthreads();
    for (int k = tid; k < n * C; k += CCN_THREADS) {
        const int ch = k % C, x = k / C;
        float sbc = 0.f, sac = 0.f;
        for (int y = 0; y < n; ++y) {
            sbc += s.Sc[(x * n + y) * C + ch];
            sac += s.Sc[(y * n + x) * C + ch];
        }
        s.Sbc[k] = sbc;
        s.Sac[k] = sac;
    }
    __syncthreads();
    for (int ch = tid; ch < C; ch += CCN_THREADS) {
        float sall = 0.f, sd = 0.f;
        for (int a = 0; a < n; ++a) {
            sall += s.Sbc[a * C + ch];
            sd += s.D1[(a * n + a) * C + ch];
        }
        s.Sall[ch] = sall;
        s.Sd[ch] = sd;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(CCN_THREADS)
ccn2_fwd_kernel(int V, const int* __restrict__ nbr_ptr, const int* __restrict__ nbr,
                const long long* __restrict__ f_off, const float* __restrict__ Fprev, int C,
                const float* __restrict__ W, const float* __restrict__ bias, int H,
                float* __restrict__ Fnext, int nmax) {
    extern __shared__ __align__(16) float smem[];
    CcnSmem s = carve(smem, nmax, C);
    const int tid = threadIdx.x;
    const int Cin = 18 * C;
    for (int i = blockIdx.x; i < V; i += gridDim.x) {
        const int n = nbr_ptr[i + 1] - nbr_ptr[i];
        __syncthreads();
        ccn2_partial_sums(s, i, nbr_ptr, nbr, f_off, Fprev, C, n);
        const float fn = (float)n;
        float* out = Fnext + f_off[i] * H;
        for (int k = tid; k < n * n * H; k += CCN_THREADS) {
            const int o = k % H, xy = k / H;
            const int x = xy / n, y = xy - x * n;
            const float* w = W + (size_t)o * Cin;
            float acc = bias ? __ldg(bias + o) : 0.f;
            for (int ch = 0; ch < C; ++ch) {
                const float sc = s.Sc[xy * C + ch];
                float wsum = 0.f;   // blocks 7..15 (index 6..14) are the same tensor: n * Sc
#pragma unroll
                for (int q = 6; q < 15; ++q) wsum += __ldg(w + q * C + ch);
                acc += (fn * (__ldg(w + ch) + wsum) + __ldg(w + 5 * C + ch)) * sc;
                acc += __ldg(w + 1 * C + ch) * s.Sbc[x * C + ch];
                acc += fn * __ldg(w + 2 * C + ch) * s.Sa[xy * C + ch];
                acc += __ldg(w + 3 * C + ch) * s.Sac[x * C + ch];
                acc += __ldg(w + 15 * C + ch) * s.D1[xy * C + ch];
                acc += __ldg(w + 16 * C + ch) * s.D2[xy * C + ch];
                if (x == y) acc += __ldg(w + 4 * C + ch) * s.Sall[ch] + __ldg(w + 17 * C + ch) * s.Sd[ch];
            }
            out[k] = fmaxf(acc, 0.f);
        }
    }
}

static int ccn_max_degree_smem(size_t floats) { return floats * sizeof(float) <= CCN_MAX_SMEM; }

extern "C" int hgnn_ccn2_update_fwd(int V, int nmax, const int* nbr_ptr, const int* nbr,
                                    const long long* f_off, const float* Fprev, int C,
                                    const float* W, const float* b, int H, float* Fnext,
                                    hgnn_stream_t stream) {
    HGNN_REQUIRE(V >= 0 && nmax >= 1 && nbr_ptr && nbr && f_off && Fprev && W && Fnext && C >= 1 && H >= 1,
                 "bad argument");
    if (V == 0) return HGNN_OK;
    size_t floats = ccn2_smem_floats(nmax, C);
    if (!ccn_max_degree_smem(floats)) {
        hgnn_set_error("hgnn_ccn2_update_fwd: receptive field %d x %d channels exceeds shared memory", nmax, C);
        return HGNN_ERR_ARG;
    }
    size_t smem = floats * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ccn2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CCN_MAX_SMEM);
        attr_set = true;
    }
    int grid = min(V, HGNN_MAX_GRID * 4);
    ccn2_fwd_kernel<<<grid, CCN_THREADS, smem, to_stream(stream)>>>(V, nbr_ptr, nbr, f_off, Fprev, C, W, b,
                                                                    H, Fnext, nmax);
    return hgnn_check_launch("hgnn_ccn2_update_fwd");
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CCN_THREADS)
ccn2_bwd_kernel(int V, const int* __restrict__ nbr_ptr, const int* __restrict__ nbr,
                const long long* __restrict__ f_off, const float* __restrict__ Fprev, int C,
                const float* __restrict__ W, int H, const float* __restrict__ Fnext,
                const float* __restrict__ gFnext, float* __restrict__ gFprev,
                float* dW, float* db, unsigned int* counter, double* accum, int nmax) {
    extern __shared__ __align__(16) float smem[];
    CcnSmem s = carve(smem, nmax, C);
    const int Cin = 18 * C;
    const int P = H * Cin + H;
    float* dacc = s.extra;                // [P]   per-CTA dW / db accumulators
    float* g = dacc + P;                  // [nmax*nmax*H] masked output gradient of a vertex
    float* RS = g + nmax * nmax * H;      // [nmax*H] row sums of g
    float* DG = RS + nmax * H;            // [H]      trace of g
    int* posm = reinterpret_cast<int*>(DG + H);   // [nmax]
    const int tid = threadIdx.x;
    for (int k = tid; k < P; k += CCN_THREADS) dacc[k] = 0.f;

    for (int v = blockIdx.x; v < V; v += gridDim.x) {
        const int n = nbr_ptr[v + 1] - nbr_ptr[v];
        __syncthreads();
        // ---------------- part A: weight gradients from vertex v's own contraction blocks
        ccn2_partial_sums(s, v, nbr_ptr, nbr, f_off, Fprev, C, n);
        {
            const float* go = gFnext + f_off[v] * H;
            const float* fo = Fnext + f_off[v] * H;
            for (int k = tid; k < n * n * H; k += CCN_THREADS) g[k] = (fo[k] > 0.f) ? go[k] : 0.f;
            __syncthreads();
            for (int k = tid; k < n * H; k += CCN_THREADS) {
                const int o = k % H, x = k / H;
                float a = 0.f;
                for (int y = 0; y < n; ++y) a += g[(x * n + y) * H + o];
                RS[k] = a;
            }
            for (int o = tid; o < H; o += CCN_THREADS) {
                float a = 0.f;
                for (int x = 0; x < n; ++x) a += g[(x * n + x) * H + o];
                DG[o] = a;
            }
            __syncthreads();
            const float fn = (float)n;
            // slots: kind in {Q1(Sc), Q3(Sa), Q16(D1), Q17(D2), Q2(Sbc), Q4(Sac), Q5/Q18, bias}
            for (int k = tid; k < 7 * H * C + H; k += CCN_THREADS) {
                if (k >= 7 * H * C) {       // bias
                    const int o = k - 7 * H * C;
                    float a = 0.f;
                    for (int x = 0; x < n; ++x) a += RS[x * H + o];
                    dacc[H * Cin + o] += a;
                    continue;
                }
                const int kind = k / (H * C), oc = k - kind * (H * C);
                const int o = oc / C, ch = oc - o * C;
                float* dw = dacc + (size_t)o * Cin;
                if (kind < 4) {
                    const float* S = kind == 0 ? s.Sc : kind == 1 ? s.Sa : kind == 2 ? s.D1 : s.D2;
                    float a = 0.f;
                    for (int xy = 0; xy < n * n; ++xy) a += g[xy * H + o] * S[xy * C + ch];
                    if (kind == 0) {
                        dw[ch] += fn * a;
                        dw[5 * C + ch] += a;
                        for (int q = 6; q < 15; ++q) dw[q * C + ch] += fn * a;
                    } else if (kind == 1) {
                        dw[2 * C + ch] += fn * a;
                    } else if (kind == 2) {
                        dw[15 * C + ch] += a;
                    } else {
                        dw[16 * C + ch] += a;
                    }
                } else if (kind < 6) {
                    const float* S = kind == 4 ? s.Sbc : s.Sac;
                    float a = 0.f;
                    for (int x = 0; x < n; ++x) a += RS[x * H + o] * S[x * C + ch];
                    dw[(kind == 4 ? 1 : 3) * C + ch] += a;
                } else {
                    dw[4 * C + ch] += DG[o] * s.Sall[ch];
                    dw[17 * C + ch] += DG[o] * s.Sd[ch];
                }
            }
        }
        // ---------------- part B: gFprev[v][p,q,:] gathered from every i in nbr(v)
        if (gFprev) {
            float* gout = gFprev + f_off[v] * C;
            // s.r holds nbr(v); accumulate in registers per owned (p,q,ch) slot: up to 4 sweeps
            for (int k0 = 0; k0 < n * n * C; k0 += CCN_THREADS) {
                const int k = k0 + tid;
                const bool act = k < n * n * C;
                const int ch = act ? k % C : 0, pq = act ? k / C : 0;
                const int p = pq / n, q = pq - p * n;
                float acc = 0.f;
                for (int ii = 0; ii < n; ++ii) {
                    const int i = s.r[ii];
                    const int ni = nbr_ptr[i + 1] - nbr_ptr[i];
                    const int ib = nbr_ptr[i];
                    __syncthreads();
                    // stage vertex i's masked gradient, row sums, trace and position map
                    {
                        const float* go = gFnext + f_off[i] * H;
                        const float* fo = Fnext + f_off[i] * H;
                        for (int t = tid; t < ni * ni * H; t += CCN_THREADS) g[t] = (fo[t] > 0.f) ? go[t] : 0.f;
                        for (int t = tid; t < n; t += CCN_THREADS) posm[t] = find_pos(nbr, ib, ib + ni, s.r[t]);
                    }
                    __syncthreads();
                    for (int t = tid; t < ni * H; t += CCN_THREADS) {
                        const int o = t % H, x = t / H;
                        float a = 0.f;
                        for (int y = 0; y < ni; ++y) a += g[(x * ni + y) * H + o];
                        RS[t] = a;
                    }
                    for (int o = tid; o < H; o += CCN_THREADS) {
                        float a = 0.f;
                        for (int x = 0; x < ni; ++x) a += g[(x * ni + x) * H + o];
                        DG[o] = a;
                    }
                    __syncthreads();
                    const int a = find_pos(nbr, ib, ib + ni, v);
                    if (act && a >= 0) {
                        const int b = posm[p], c = posm[q];
                        if (b >= 0 && c >= 0) {
                            const float fni = (float)ni;
                            for (int o = 0; o < H; ++o) {
                                const float* w = W + (size_t)o * Cin;
                                float wsum = 0.f;
#pragma unroll
                                for (int qq = 6; qq < 15; ++qq) wsum += __ldg(w + qq * C + ch);
                                const float walpha = fni * (__ldg(w + ch) + wsum) + __ldg(w + 5 * C + ch);
                                const float gab = g[(a * ni + b) * H + o];
                                float t = walpha * gab;
                                t += __ldg(w + 1 * C + ch) * RS[a * H + o];
                                t += fni * __ldg(w + 2 * C + ch) * g[(b * ni + c) * H + o];
                                t += __ldg(w + 3 * C + ch) * RS[b * H + o];
                                t += __ldg(w + 4 * C + ch) * DG[o];
                                if (b == c) t += __ldg(w + 15 * C + ch) * gab;
                                if (a == c) t += __ldg(w + 16 * C + ch) * g[(b * ni + a) * H + o];
                                if (a == b && b == c) t += __ldg(w + 17 * C + ch) * DG[o];
                                acc += t;
                            }
                        }
                    }
                }
                if (act) gout[k] = acc;
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < P; k += CCN_THREADS) accum_add(accum, P, hgnn_ws_bins(P), k, (double)dacc[k]);
    if (last_block_ticket(counter)) {
        for (int k = tid; k < P; k += CCN_THREADS) {
            const float a = (float)accum_take(accum, P, hgnn_ws_bins(P), k);
            if (k < H * Cin) dW[k] = a; else db[k - H * Cin] = a;
        }
        if (tid == 0) *counter = 0;
    }
}

extern "C" int hgnn_ccn2_update_bwd(int V, int nmax, const int* nbr_ptr, const int* nbr,
                                    const long long* f_off, const float* Fprev, int C,
                                    const float* W, int H, const float* Fnext, const float* gFnext,
                                    float* gFprev, float* dW, float* db, void* ws, long long ws_bytes,
                                    hgnn_stream_t stream) {
    HGNN_REQUIRE(V >= 0 && nmax >= 1 && nbr_ptr && nbr && f_off && Fprev && W && Fnext && gFnext && dW && db && ws,
                 "bad argument");
    const int P = H * 18 * C + H;
    if (ws_bytes < hgnn_workspace_bytes(P)) {
        hgnn_set_error("hgnn_ccn2_update_bwd: workspace too small");
        return HGNN_ERR_WORKSPACE;
    }
    if (V == 0) return HGNN_OK;
    size_t floats = ccn2_smem_floats(nmax, C) + P + (size_t)nmax * nmax * H + (size_t)nmax * H + H + nmax;
    if (!ccn_max_degree_smem(floats)) {
        hgnn_set_error("hgnn_ccn2_update_bwd: receptive field %d exceeds shared memory", nmax);
        return HGNN_ERR_ARG;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ccn2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CCN_MAX_SMEM);
        attr_set = true;
    }
    int grid = min(V, HGNN_SM_COUNT * 8);
    ccn2_bwd_kernel<<<grid, CCN_THREADS, floats * sizeof(float), to_stream(stream)>>>(
        V, nbr_ptr, nbr, f_off, Fprev, C, W, H, Fnext, gFnext, gFprev, dW, db, (unsigned int*)ws,
        (double*)((char*)ws + HGNN_WS_HEADER), nmax);
    return hgnn_check_launch("hgnn_ccn2_update_bwd");
}

// ---------------------------------------------------------------------------------------------
// 1-D variant (functions/utils_ccn.py:303-324): F[v] is (d_v, C) at row offset nbr_ptr[v]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CCN_THREADS)
ccn1_fwd_kernel(int V, const int* __restrict__ nbr_ptr, const int* __restrict__ nbr,
                const float* __restrict__ Fprev, int C, const float* __restrict__ W,
                const float* __restrict__ bias, int H, float* __restrict__ Fnext, int nmax) {
    extern __shared__ __align__(16) float smem[];
    int* r = reinterpret_cast<int*>(smem);
    int* m = r + nmax;
    float* rowc = reinterpret_cast<float*>(m + nmax * nmax);   // [n*C]  sum_a T[a,x]
    float* colc = rowc + nmax * C;                             // [n*C]  sum_b T[x,b]
    const int tid = threadIdx.x;
    for (int i = blockIdx.x; i < V; i += gridDim.x) {
        const int base = nbr_ptr[i], n = nbr_ptr[i + 1] - base;
        __syncthreads();
        for (int k = tid; k < n; k += CCN_THREADS) r[k] = nbr[base + k];
        __syncthreads();
        for (int k = tid; k < n * n; k += CCN_THREADS) {
            const int a = k / n, b = k - a * n;
            m[k] = find_pos(nbr, nbr_ptr[r[a]], nbr_ptr[r[a] + 1], r[b]);
        }
        __syncthreads();
        for (int k = tid; k < n * C; k += CCN_THREADS) {
            const int ch = k % C, x = k / C;
            float rs = 0.f, cs = 0.f;
            for (int t = 0; t < n; ++t) {
                const int p1 = m[t * n + x];   // T[t, x] = F_{r_t}[pos(r_x)]
                if (p1 >= 0) rs += __ldg(Fprev + ((size_t)nbr_ptr[r[t]] + p1) * C + ch);
                const int p2 = m[x * n + t];   // T[x, t] = F_{r_x}[pos(r_t)]
                if (p2 >= 0) cs += __ldg(Fprev + ((size_t)nbr_ptr[r[x]] + p2) * C + ch);
            }
            rowc[k] = rs;
            colc[k] = cs;
        }
        __syncthreads();
        for (int k = tid; k < n * H; k += CCN_THREADS) {
            const int o = k % H, x = k / H;
            const float* w = W + (size_t)o * 2 * C;
            float acc = bias ? __ldg(bias + o) : 0.f;
            for (int ch = 0; ch < C; ++ch)
                acc += __ldg(w + ch) * rowc[x * C + ch] + __ldg(w + C + ch) * colc[x * C + ch];
            Fnext[((size_t)base + x) * H + o] = fmaxf(acc, 0.f);
        }
    }
}

extern "C" int hgnn_ccn1_update_fwd(int V, int nmax, const int* nbr_ptr, const int* nbr,
                                    const float* Fprev, int C, const float* W, const float* b, int H,
                                    float* Fnext, hgnn_stream_t stream) {
    HGNN_REQUIRE(V >= 0 && nmax >= 1 && nbr_ptr && nbr && Fprev && W && Fnext && C >= 1 && H >= 1, "bad argument");
    if (V == 0) return HGNN_OK;
    size_t floats = (size_t)nmax + (size_t)nmax * nmax + 2 * (size_t)nmax * C;
    if (!ccn_max_degree_smem(floats)) {
        hgnn_set_error("hgnn_ccn1_update_fwd: receptive field %d exceeds shared memory", nmax);
        return HGNN_ERR_ARG;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ccn1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CCN_MAX_SMEM);
        attr_set = true;
    }
    ccn1_fwd_kernel<<<min(V, HGNN_MAX_GRID * 4), CCN_THREADS, floats * sizeof(float), to_stream(stream)>>>(
        V, nbr_ptr, nbr, Fprev, C, W, b, H, Fnext, nmax);
    return hgnn_check_launch("hgnn_ccn1_update_fwd");
}

__global__ void __launch_bounds__(CCN_THREADS)
ccn1_bwd_kernel(int V, const int* __restrict__ nbr_ptr, const int* __restrict__ nbr,
                const float* __restrict__ Fprev, int C, const float* __restrict__ W, int H,
                const float* __restrict__ Fnext, const float* __restrict__ gFnext,
                float* __restrict__ gFprev, float* dW, float* db, unsigned int* counter,
                double* accum, int nmax) {
    extern __shared__ __align__(16) float smem[];
    int* r = reinterpret_cast<int*>(smem);
    int* m = r + nmax;
    float* rowc = reinterpret_cast<float*>(m + nmax * nmax);
    float* colc = rowc + nmax * C;
    const int P = H * 2 * C + H;
    float* dacc = colc + nmax * C;     // [P]
    float* g = dacc + P;               // [nmax*H]
    const int tid = threadIdx.x;
    for (int k = tid; k < P; k += CCN_THREADS) dacc[k] = 0.f;
    for (int v = blockIdx.x; v < V; v += gridDim.x) {
        const int base = nbr_ptr[v], n = nbr_ptr[v + 1] - base;
        __syncthreads();
        for (int k = tid; k < n; k += CCN_THREADS) r[k] = nbr[base + k];
        __syncthreads();
        for (int k = tid; k < n * n; k += CCN_THREADS) {
            const int a = k / n, b = k - a * n;
            m[k] = find_pos(nbr, nbr_ptr[r[a]], nbr_ptr[r[a] + 1], r[b]);
        }
        for (int k = tid; k < n * H; k += CCN_THREADS) {
            const size_t idx = (size_t)base * H + k;
            g[k] = (Fnext[idx] > 0.f) ? gFnext[idx] : 0.f;
        }
        __syncthreads();
        for (int k = tid; k < n * C; k += CCN_THREADS) {
            const int ch = k % C, x = k / C;
            float rs = 0.f, cs = 0.f;
            for (int t = 0; t < n; ++t) {
                const int p1 = m[t * n + x];
                if (p1 >= 0) rs += __ldg(Fprev + ((size_t)nbr_ptr[r[t]] + p1) * C + ch);
                const int p2 = m[x * n + t];
                if (p2 >= 0) cs += __ldg(Fprev + ((size_t)nbr_ptr[r[x]] + p2) * C + ch);
            }
            rowc[k] = rs;
            colc[k] = cs;
        }
        __syncthreads();
        for (int k = tid; k < P; k += CCN_THREADS) {
            float a = 0.f;
            if (k < H * 2 * C) {
                const int o = k / (2 * C), cc = k - o * 2 * C;
                const float* S = cc < C ? rowc : colc;
                const int ch = cc < C ? cc : cc - C;
                for (int x = 0; x < n; ++x) a += g[x * H + o] * S[x * C + ch];
            } else {
                const int o = k - H * 2 * C;
                for (int x = 0; x < n; ++x) a += g[x * H + o];
            }
            dacc[k] += a;
        }
        // gFprev[v][p, ch] = sum_{i in nbr(v)} gT_i[a = pos_i(v), b = pos_i(nbr_v[p])]
        //   gT_i[a, b, ch] = sum_o W[o, ch] g_i[b, o] + W[o, C+ch] g_i[a, o]
        if (gFprev) {
            for (int k = tid; k < n * C; k += CCN_THREADS) {
                const int ch = k % C, p = k / C;
                float acc = 0.f;
                for (int ii = 0; ii < n; ++ii) {
                    const int i = r[ii];
                    const int ib = nbr_ptr[i], ni = nbr_ptr[i + 1] - ib;
                    const int a = find_pos(nbr, ib, ib + ni, v);
                    const int b = find_pos(nbr, ib, ib + ni, r[p]);
                    if (a < 0 || b < 0) continue;
                    for (int o = 0; o < H; ++o) {
                        const size_t ia = ((size_t)ib + a) * H + o, ibb = ((size_t)ib + b) * H + o;
                        const float ga = (Fnext[ia] > 0.f) ? gFnext[ia] : 0.f;
                        const float gb = (Fnext[ibb] > 0.f) ? gFnext[ibb] : 0.f;
                        acc += __ldg(W + (size_t)o * 2 * C + ch) * gb + __ldg(W + (size_t)o * 2 * C + C + ch) * ga;
                    }
                }
                gFprev[((size_t)base + p) * C + ch] = acc;
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < P; k += CCN_THREADS) accum_add(accum, P, hgnn_ws_bins(P), k, (double)dacc[k]);
    if (last_block_ticket(counter)) {
        for (int k = tid; k < P; k += CCN_THREADS) {
            const float a = (float)accum_take(accum, P, hgnn_ws_bins(P), k);
            if (k < H * 2 * C) dW[k] = a; else db[k - H * 2 * C] = a;
        }
        if (tid == 0) *counter = 0;
    }
}

extern "C" int hgnn_ccn1_update_bwd(int V, int nmax, const int* nbr_ptr, const int* nbr,
                                    const float* Fprev, int C, const float* W, int H,
                                    const float* Fnext, const float* gFnext, float* gFprev, float* dW,
                                    float* db, void* ws, long long ws_bytes, hgnn_stream_t stream) {
    HGNN_REQUIRE(V >= 0 && nmax >= 1 && nbr_ptr && nbr && Fprev && W && Fnext && gFnext && dW && db && ws,
                 "bad argument");
    const int P = H * 2 * C + H;
    if (ws_bytes < hgnn_workspace_bytes(P)) {
        hgnn_set_error("hgnn_ccn1_update_bwd: workspace too small");
        return HGNN_ERR_WORKSPACE;
    }
    if (V == 0) return HGNN_OK;
    size_t floats = (size_t)nmax + (size_t)nmax * nmax + 2 * (size_t)nmax * C + P + (size_t)nmax * H;
    if (!ccn_max_degree_smem(floats)) {
        hgnn_set_error("hgnn_ccn1_update_bwd: receptive field %d exceeds shared memory", nmax);
        return HGNN_ERR_ARG;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ccn1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CCN_MAX_SMEM);
        attr_set = true;
    }
    int grid = min(V, HGNN_SM_COUNT * 8);
    ccn1_bwd_kernel<<<grid, CCN_THREADS, floats * sizeof(float), to_stream(stream)>>>(
        V, nbr_ptr, nbr, Fprev, C, W, H, Fnext, gFnext, gFprev, dW, db, (unsigned int*)ws,
        (double*)((char*)ws + HGNN_WS_HEADER), nmax);
    return hgnn_check_launch("hgnn_ccn1_update_bwd");
}

// ---------------------------------------------------------------------------------------------
// stand-alone collapse6to3 on a general rank-6 tensor (functions/contraction.py:106-121), fwd + bwd
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t idx5(int n, int a, int b, int c, int d, int e) {
    return ((((size_t)a * n + b) * n + c) * n + d) * n + e;
}

__global__ void collapse6to3_kernel(const float* __restrict__ F6, int C, int n, float* __restrict__ out) {
    const long long total = (long long)n * n * 18 * C;
    const size_t n5 = (size_t)n * n * n * n * n;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(t % C);
        const int k = (int)((t / C) % 18);
        const int y = (int)((t / (18 * C)) % n);
        const int x = (int)(t / ((long long)18 * C * n));
        const float* G = F6 + (size_t)ch * n5;
        float acc = 0.f;
        if (k < 5) {            // contraction.py:51-57: fix two axes, sum the other three
            for (int u = 0; u < n; ++u)
                for (int v = 0; v < n; ++v)
                    for (int w = 0; w < n; ++w) {
                        size_t id;
                        if (k == 0) id = idx5(n, x, y, u, v, w);
                        else if (k == 1) id = idx5(n, x, u, v, y, w);
                        else if (k == 2) id = idx5(n, u, x, y, v, w);
                        else if (k == 3) id = idx5(n, u, x, v, y, w);
                        else id = idx5(n, u, v, w, x, y);
                        acc += G[id];
                    }
        } else if (k == 5) {    // :70  out[a,b] = sum_{c,e} G[a,b,c,c,e]
            for (int c = 0; c < n; ++c)
                for (int e = 0; e < n; ++e) acc += G[idx5(n, x, y, c, c, e)];
        } else if (k < 15) {    // :71-80 (identity permutation nine times): sum_{c,d} G[a,b,c,d,d]
            for (int c = 0; c < n; ++c)
                for (int d = 0; d < n; ++d) acc += G[idx5(n, x, y, c, d, d)];
        } else if (k == 15) {   // :97  out[a,d] = sum_b G[a,b,b,d,b]
            for (int b = 0; b < n; ++b) acc += G[idx5(n, x, b, b, y, b)];
        } else if (k == 16) {   // :98  out[b,d] = sum_a G[a,b,a,d,a]
            for (int a = 0; a < n; ++a) acc += G[idx5(n, a, x, a, y, a)];
        } else {                // :99  out[d,e] = sum_a G[a,a,a,d,e]
            for (int a = 0; a < n; ++a) acc += G[idx5(n, a, a, a, x, y)];
        }
        out[t] = acc;
    }
}

extern "C" int hgnn_ccn2_collapse6to3(const float* F6, int C, int n, float* out, hgnn_stream_t stream) {
    HGNN_REQUIRE(F6 && out && C >= 1 && n >= 1, "bad argument");
    long long total = (long long)n * n * 18 * C;
    collapse6to3_kernel<<<min(ceil_div(total, 128), HGNN_MAX_GRID), 128, 0, to_stream(stream)>>>(F6, C, n, out);
    return hgnn_check_launch("hgnn_ccn2_collapse6to3");
}

__global__ void collapse6to3_bwd_kernel(const float* __restrict__ gout, int C, int n, float* __restrict__ gF6) {
    const size_t n5 = (size_t)n * n * n * n * n;
    const long long total = (long long)C * n5;
    const int ld = 18 * C;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(t / n5);
        size_t rem = (size_t)(t % n5);
        const int e = (int)(rem % n); rem /= n;
        const int d = (int)(rem % n); rem /= n;
        const int c = (int)(rem % n); rem /= n;
        const int b = (int)(rem % n);
        const int a = (int)(rem / n);
#define GO(x, y, k) gout[((size_t)(x) * n + (y)) * ld + (k) * C + ch]
        float acc = GO(a, b, 0) + GO(a, d, 1) + GO(b, c, 2) + GO(b, d, 3) + GO(d, e, 4);
        if (c == d) acc += GO(a, b, 5);
        if (d == e) {
            float s9 = 0.f;
            for (int k = 6; k < 15; ++k) s9 += GO(a, b, k);
            acc += s9;
        }
        if (b == c && c == e) acc += GO(a, d, 15);
        if (a == c && c == e) acc += GO(b, d, 16);
        if (a == b && b == c) acc += GO(d, e, 17);
#undef GO
        gF6[t] = acc;
    }
}

extern "C" int hgnn_ccn2_collapse6to3_bwd(const float* gout, int C, int n, float* gF6, hgnn_stream_t stream) {
    HGNN_REQUIRE(gout && gF6 && C >= 1 && n >= 1, "bad argument");
    long long total = (long long)C * n * n * n * n * n;
    collapse6to3_bwd_kernel<<<min(ceil_div(total, 256), HGNN_MAX_GRID), 256, 0, to_stream(stream)>>>(gout, C, n, gF6);
    return hgnn_check_launch("hgnn_ccn2_collapse6to3_bwd");
}
