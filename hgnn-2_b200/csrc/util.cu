// util.cu -- error handling, layout conversion, dense<->CSR, scan, segment sums.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void hgnn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int hgnn_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        hgnn_set_error("%s: CUDA error: %s", what, cudaGetErrorString(e));
        return HGNN_ERR_CUDA;
    }
    return HGNN_OK;
}

extern "C" const char* hgnn_last_error(void) { return g_err; }
extern "C" int hgnn_version(void) { return HGNN_B200_VERSION; }

// Grid cap shared by the workspace query and every kernel that reduces across CTAs.
int hgnn_grid_cap(int width) {
    long long budget = 32ll << 20;
    long long g = budget / ((long long)(width < 1 ? 1 : width) * 8);
    if (g >= HGNN_MAX_GRID) return HGNN_MAX_GRID;
    g = (g / HGNN_SM_COUNT) * HGNN_SM_COUNT;
    return (int)(g < HGNN_SM_COUNT ? HGNN_SM_COUNT : g);
}

// Workspace = 256-byte header (ticket counter) + hgnn_ws_bins(width) x `width` fp64 accumulators; it must be all zero on
// entry and every kernel leaves it all zero again.
extern "C" long long hgnn_workspace_bytes(int width) {
    const int w = width < 1 ? 1 : width;
    return HGNN_WS_HEADER + (long long)hgnn_ws_bins(w) * w * 8;
}

extern "C" int hgnn_bins_for(int width) { return hgnn_ws_bins(width < 1 ? 1 : width); }

// ---------------------------------------------------------------------------------------------
// layout: (bs, F, Nmax) channel-major padded <-> packed (R, F)
// One CTA per (graph, 32-node tile); transposes through shared memory so both sides coalesce.
// ---------------------------------------------------------------------------------------------
__global__ void pack_rows_kernel(const float* __restrict__ dense, int F, int Nmax,
                                 const int* __restrict__ off, float* __restrict__ packed) {
    __shared__ float tile[32][33];
    const int b = blockIdx.y;
    const int r0 = off[b], n = off[b + 1] - r0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int v0 = blockIdx.x * 32; v0 < n; v0 += gridDim.x * 32) {
        for (int f0 = 0; f0 < F; f0 += 32) {
            for (int i = ty; i < 32; i += 8) {
                int f = f0 + i, v = v0 + tx;
                tile[i][tx] = (f < F && v < n) ? dense[((size_t)b * F + f) * Nmax + v] : 0.f;
            }
            __syncthreads();
            for (int i = ty; i < 32; i += 8) {
                int v = v0 + i, f = f0 + tx;
                if (v < n && f < F) packed[(size_t)(r0 + v) * F + f] = tile[tx][i];
            }
            __syncthreads();
        }
    }
}

__global__ void unpack_rows_kernel(const float* __restrict__ packed, int F, int Nmax,
                                   const int* __restrict__ off, const float* __restrict__ pad_fill,
                                   float* __restrict__ dense) {
    __shared__ float tile[32][33];
    const int b = blockIdx.y;
    const int r0 = off[b], n = off[b + 1] - r0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int v0 = blockIdx.x * 32; v0 < Nmax; v0 += gridDim.x * 32) {
        for (int f0 = 0; f0 < F; f0 += 32) {
            for (int i = ty; i < 32; i += 8) {
                int v = v0 + i, f = f0 + tx;
                float x = 0.f;
                if (f < F) x = (v < n) ? packed[(size_t)(r0 + v) * F + f] : (pad_fill ? pad_fill[f] : 0.f);
                tile[i][tx] = x;
            }
            __syncthreads();
            for (int i = ty; i < 32; i += 8) {
                int f = f0 + i, v = v0 + tx;
                if (f < F && v < Nmax) dense[((size_t)b * F + f) * Nmax + v] = tile[tx][i];
            }
            __syncthreads();
        }
    }
}

extern "C" int hgnn_pack_rows(const float* dense, int bs, int F, int Nmax, const int* off,
                              float* packed, hgnn_stream_t stream) {
    HGNN_REQUIRE(bs >= 0 && F >= 0 && Nmax >= 0, "negative size");
    if (bs == 0 || F == 0 || Nmax == 0) return HGNN_OK;
    HGNN_REQUIRE(bs <= 65535, "bs > 65535");
    dim3 grid(min(ceil_div(Nmax, 32), 1024), bs);
    pack_rows_kernel<<<grid, 256, 0, to_stream(stream)>>>(dense, F, Nmax, off, packed);
    return hgnn_check_launch("hgnn_pack_rows");
}

extern "C" int hgnn_unpack_rows(const float* packed, int bs, int F, int Nmax, const int* off,
                                const float* pad_fill, float* dense, hgnn_stream_t stream) {
    HGNN_REQUIRE(bs >= 0 && F >= 0 && Nmax >= 0, "negative size");
    if (bs == 0 || F == 0 || Nmax == 0) return HGNN_OK;
    HGNN_REQUIRE(bs <= 65535, "bs > 65535");
    dim3 grid(min(ceil_div(Nmax, 32), 1024), bs);
    unpack_rows_kernel<<<grid, 256, 0, to_stream(stream)>>>(packed, F, Nmax, off, pad_fill, dense);
    return hgnn_check_launch("hgnn_unpack_rows");
}

// ---------------------------------------------------------------------------------------------
// dense -> CSR.  One warp per (graph, row): ballot-compaction keeps columns ascending.
// ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void dense_rows_kernel(const float* __restrict__ D1, const float* __restrict__ D2,
                                  long long sb, long long sr, long long sc, int bs,
                                  const int* __restrict__ row_off, const int* __restrict__ col_off,
                                  int* __restrict__ rowcnt, const int* __restrict__ rowptr,
                                  int* __restrict__ col, float* __restrict__ val1,
                                  float* __restrict__ val2) {
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int R = row_off[bs];
    for (int row = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < R;
         row += gridDim.x * warps_per_cta) {
        // locate graph b: row_off[b] <= row < row_off[b+1]
        int lo = 0, hi = bs;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (row_off[mid] <= row) lo = mid; else hi = mid;
        }
        const int b = lo;
        const int r = row - row_off[b];
        const int c0 = col_off[b], nc = col_off[b + 1] - c0;
        const float* p1 = D1 + (long long)b * sb + (long long)r * sr;
        const float* p2 = D2 ? D2 + (long long)b * sb + (long long)r * sr : nullptr;
        int count = 0;
        int base = FILL ? rowptr[row] : 0;
        for (int cb = 0; cb < nc; cb += 32) {
            int c = cb + lane;
            float a = 0.f, d = 0.f;
            if (c < nc) {
                a = p1[(long long)c * sc];
                if (p2) d = p2[(long long)c * sc];
            }
            bool nz = (a != 0.f) || (d != 0.f);
            unsigned m = __ballot_sync(0xffffffffu, nz);
            if (FILL && nz) {
                int pos = base + count + __popc(m & ((1u << lane) - 1u));
                col[pos] = c0 + c;
                val1[pos] = a;
                if (val2) val2[pos] = d;
            }
            count += __popc(m);
        }
        if (!FILL && lane == 0) rowcnt[row] = count;
    }
}

extern "C" int hgnn_dense_count_nnz(const float* D1, const float* D2, long long sb, long long sr,
                                    long long sc, int bs, const int* row_off, const int* col_off,
                                    int* rowcnt, hgnn_stream_t stream) {
    HGNN_REQUIRE(D1 && row_off && col_off && rowcnt && bs > 0, "bad argument");
    dense_rows_kernel<false><<<HGNN_SM_COUNT * 4, 256, 0, to_stream(stream)>>>(
        D1, D2, sb, sr, sc, bs, row_off, col_off, rowcnt, nullptr, nullptr, nullptr, nullptr);
    return hgnn_check_launch("hgnn_dense_count_nnz");
}

extern "C" int hgnn_dense_fill_csr(const float* D1, const float* D2, long long sb, long long sr,
                                   long long sc, int bs, const int* row_off, const int* col_off,
                                   const int* rowptr, int* col, float* val1, float* val2,
                                   hgnn_stream_t stream) {
    HGNN_REQUIRE(D1 && row_off && col_off && rowptr && bs > 0, "bad argument");
    dense_rows_kernel<true><<<HGNN_SM_COUNT * 4, 256, 0, to_stream(stream)>>>(
        D1, D2, sb, sr, sc, bs, row_off, col_off, nullptr, rowptr, col, val1, val2);
    return hgnn_check_launch("hgnn_dense_fill_csr");
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (pack-time bookkeeping; one CTA walks the array with a running carry).
// ---------------------------------------------------------------------------------------------
__global__ void scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < n) ? in[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[w] = x;
        __syncthreads();
        if (w == 0) {
            int t = (lane < (blockDim.x >> 5)) ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[lane] = t;  // inclusive totals of warps
        }
        __syncthreads();
        int carry = carry_s;
        int prefix = carry + (w > 0 ? warp_tot[w - 1] : 0) + x - v;
        if (i < n) out[i] = prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

extern "C" int hgnn_exclusive_scan_i32(const int* in, int* out, int n, hgnn_stream_t stream) {
    HGNN_REQUIRE(n >= 0 && out, "bad argument");
    scan_kernel<<<1, 1024, 0, to_stream(stream)>>>(in, out, n);
    return hgnn_check_launch("hgnn_exclusive_scan_i32");
}

// ---------------------------------------------------------------------------------------------
// CSR -> dense, row sums
// ---------------------------------------------------------------------------------------------
__global__ void csr_to_dense_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                    const float* __restrict__ val, int bs,
                                    const int* __restrict__ row_off, const int* __restrict__ col_off,
                                    float* __restrict__ D, long long sb, long long sr, long long sc) {
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int R = row_off[bs];
    for (int row = blockIdx.x * wpc + (threadIdx.x >> 5); row < R; row += gridDim.x * wpc) {
        int lo = 0, hi = bs;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (row_off[mid] <= row) lo = mid; else hi = mid;
        }
        const int b = lo;
        float* p = D + (long long)b * sb + (long long)(row - row_off[b]) * sr;
        const int c0 = col_off[b];
        for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32)
            p[(long long)(col[k] - c0) * sc] = val[k];
    }
}

extern "C" int hgnn_csr_to_dense(const int* rowptr, const int* col, const float* val, int bs,
                                 const int* row_off, const int* col_off, float* D, long long sb,
                                 long long sr, long long sc, hgnn_stream_t stream) {
    HGNN_REQUIRE(rowptr && row_off && col_off && D && bs > 0, "bad argument");
    csr_to_dense_kernel<<<HGNN_SM_COUNT * 4, 256, 0, to_stream(stream)>>>(rowptr, col, val, bs, row_off,
                                                                         col_off, D, sb, sr, sc);
    return hgnn_check_launch("hgnn_csr_to_dense");
}

__global__ void csr_row_sums_kernel(const int* __restrict__ rowptr, const float* __restrict__ val,
                                    int R, float* __restrict__ out) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k];
        out[r] = s;
    }
}

extern "C" int hgnn_csr_row_sums(const int* rowptr, const float* val, int R, float* out,
                                 hgnn_stream_t stream) {
    HGNN_REQUIRE(R >= 0, "bad argument");
    if (R == 0) return HGNN_OK;
    csr_row_sums_kernel<<<min(ceil_div(R, 256), HGNN_MAX_GRID), 256, 0, to_stream(stream)>>>(rowptr, val, R, out);
    return hgnn_check_launch("hgnn_csr_row_sums");
}

// ---------------------------------------------------------------------------------------------
// readout: per-graph sums and their broadcast (layers_mnb.py:92, :386)
// ---------------------------------------------------------------------------------------------
__global__ void segment_sum_kernel(const float* __restrict__ Y, int F, const int* __restrict__ off,
                                   const float* __restrict__ pad_count,
                                   const float* __restrict__ bias, float* __restrict__ out) {
    // one CTA per graph; thread t -> feature t % F, row lane t / F; fixed-order tree reduce
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    const int r0 = off[b], r1 = off[b + 1];
    for (int f0 = 0; f0 < F; f0 += blockDim.x) {
        int groups = (F - f0 >= (int)blockDim.x) ? 1 : blockDim.x / (F - f0);
        int width = (F - f0 >= (int)blockDim.x) ? blockDim.x : (F - f0);
        int f = f0 + threadIdx.x % width, g = threadIdx.x / width;
        double s = 0.0;
        if (g < groups)
            for (int r = r0 + g; r < r1; r += groups) s += (double)Y[(size_t)r * F + f];
        sm[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x < width) {
            double t = 0.0;
            for (int k = 0; k < groups; ++k) t += sm[k * width + threadIdx.x];
            float extra = (pad_count && bias) ? pad_count[b] * bias[f] : 0.f;
            out[(size_t)b * F + f] = (float)t + extra;
        }
        __syncthreads();
    }
}

extern "C" int hgnn_segment_sum(const float* Y, int bs, int F, const int* off,
                                const float* pad_count, const float* bias, float* out,
                                hgnn_stream_t stream) {
    HGNN_REQUIRE(bs >= 0 && F > 0, "bad argument");
    if (bs == 0) return HGNN_OK;
    segment_sum_kernel<<<bs, 256, 256 * sizeof(double), to_stream(stream)>>>(Y, F, off, pad_count, bias, out);
    return hgnn_check_launch("hgnn_segment_sum");
}

__global__ void segment_bcast_kernel(const float* __restrict__ g, int bs, int F,
                                     const int* __restrict__ off, float* __restrict__ G) {
    const int b = blockIdx.y;
    const int r0 = off[b];
    const long long n = (long long)(off[b + 1] - r0) * F;
    float* dst = G + (size_t)r0 * F;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        dst[i] = g[(size_t)b * F + (int)(i % F)];
}

extern "C" int hgnn_segment_bcast(const float* g, int bs, int F, const int* off, float* G,
                                  hgnn_stream_t stream) {
    HGNN_REQUIRE(bs >= 0 && F > 0, "bad argument");
    if (bs == 0) return HGNN_OK;
    HGNN_REQUIRE(bs <= 65535, "bs > 65535");
    dim3 grid(32, bs);
    segment_bcast_kernel<<<grid, 256, 0, to_stream(stream)>>>(g, bs, F, off, G);
    return hgnn_check_launch("hgnn_segment_bcast");
}

// ---------------------------------------------------------------------------------------------
// block-diagonal fix-up: the host copies every graph's index arrays RAW into the staging buffer;
// this adds the per-graph row / column / nnz offsets in place after the H2D copy.
// table[e] = (array offset, length, segment-pointer offset, segment-addend offset), in 4-byte words
// from `base`;  arr[i] += addend[g] for i in [segptr[g], segptr[g+1]).
// ---------------------------------------------------------------------------------------------
__global__ void fixup_offsets_kernel(int* __restrict__ base, const int* __restrict__ table, int n_seg) {
    const int e = blockIdx.z, g = blockIdx.y;
    const int4 t = *reinterpret_cast<const int4*>(table + 4 * e);
    int* arr = base + t.x;
    const int* segp = base + t.z;
    const int add = base[t.w + g];
    const int lo = segp[g], hi = min(segp[g + 1], t.y);
    if (add == 0) return;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) arr[i] += add;
}

extern "C" int hgnn_fixup_offsets(int* base, const int* table, int n_entries, int n_seg, hgnn_stream_t stream) {
    HGNN_REQUIRE(base && table && n_entries >= 0 && n_seg >= 0, "bad argument");
    if (n_entries == 0 || n_seg == 0) return HGNN_OK;
    HGNN_REQUIRE(n_seg <= 65535 && n_entries <= 65535, "too many segments");
    dim3 grid(8, n_seg, n_entries);
    fixup_offsets_kernel<<<grid, 256, 0, to_stream(stream)>>>(base, table, n_seg);
    return hgnn_check_launch("hgnn_fixup_offsets");
}
