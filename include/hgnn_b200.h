/* hgnn_b200.h -- C ABI of the B200-native (sm_100a) HGNN-2 aggregation hot path.
 *
 * The reference (AmmieQi/HGNN-2) has no FFI: its boundary is Python.  Each
 * entry point below names the reference code it replaces (paths relative to the reference
 * checkout).  The Python host layer (hgnn-2_b200/) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates,
 *     never frees, never synchronises; the caller (PyTorch caching allocator) owns all buffers;
 *   - return value 0 = ok, <0 = error; hgnn_last_error() gives the thread-local message;
 *   - feature matrices are row-major "packed rows": (R, F) fp32, one row per real node / line-graph
 *     node of the block-diagonal batch (no padding).  The reference's (bs, F, Nmax) channel-major
 *     padded layout is converted at the boundary by hgnn_pack_rows / hgnn_unpack_rows;
 *   - indices are int32; CSR column indices are global row numbers of the batch.
 */
#ifndef HGNN_B200_H
#define HGNN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HGNN_B200_VERSION 100

#define HGNN_OP_IDENT 0 /* y = x                         (W[:,:,0] = I,  functions/operators.py:20) */
#define HGNN_OP_DIAG 1  /* y = diag[r] * x               (W[:,:,1] = D,  functions/operators.py:22-23) */
#define HGNN_OP_CSR 2   /* y = sum_k val[k] x[col[k]]    (A^(2^j), AL^(2^j), Pm, Pd and transposes) */
#define HGNN_MAX_OPS 8

#define HGNN_OK 0
#define HGNN_ERR_ARG (-1)
#define HGNN_ERR_CUDA (-2)
#define HGNN_ERR_WORKSPACE (-3)

typedef void* hgnn_stream_t;

/* One graph operator applied to packed rows. */
typedef struct hgnn_op_t {
    int kind;            /* HGNN_OP_* */
    const float* diag;   /* DIAG: (R,) */
    const int* rowptr;   /* CSR: (R+1,) */
    const int* col;      /* CSR: (nnz,) */
    const float* val;    /* CSR: (nnz,) */
    /* optional run-length part (engine kernels only; NULL = none): row r additionally gets
     * sum_{k in [rng_rowptr[r], rng_rowptr[r+1])} rng_val[k] * sum_{c in [rng_lo[id], rng_hi[id])} x[c],
     * id = rng_id[k] - long runs of equal consecutive entries stored once (sparse_ops.split_runs). */
    const int* rng_rowptr;
    const int* rng_id;
    const float* rng_val;
    const int* rng_lo;
    const int* rng_hi;
    long long nnz;       /* CSR: number of stored entries, 0 = unknown (a scheduling hint only) */
    int rng_n;           /* number of distinct ranges (length of rng_lo / rng_hi), 0 = unknown */
} hgnn_op_t;

const char* hgnn_last_error(void);
int hgnn_version(void);
/* Workspace (bytes) for cross-CTA reductions: a ticket counter + binned fp64 accumulators for `width` values (see each
 * call for its width).  The buffer must be ALL ZERO before the first use; every kernel leaves it all
 * zero again, so one buffer per stream can serve every call. */
long long hgnn_workspace_bytes(int width);

/* ---- layout conversion (boundary of functions/batching.py:77-185) ------------------------- */
/* dense (bs, F, Nmax) channel-major padded  ->  packed (R, F); off = (bs+1) row offsets. */
int hgnn_pack_rows(const float* dense, int bs, int F, int Nmax, const int* off, float* packed,
                   hgnn_stream_t stream);
/* packed (R, F) -> dense (bs, F, Nmax); padded slots get pad_fill[f] (NULL = 0). */
int hgnn_unpack_rows(const float* packed, int bs, int F, int Nmax, const int* off,
                     const float* pad_fill, float* dense, hgnn_stream_t stream);

/* ---- dense -> CSR (accepting the reference's dense W / WL / Pm / Pd tensors) --------------- */
/* D[b][r][c] = base[b*sb + r*sr + c*sc] (element strides); rows r < row_off[b+1]-row_off[b],
 * cols c < col_off[b+1]-col_off[b].  Pattern = (D1 != 0) | (D2 != 0); D2 may be NULL. */
int hgnn_dense_count_nnz(const float* D1, const float* D2, long long sb, long long sr, long long sc,
                         int bs, const int* row_off, const int* col_off, int* rowcnt,
                         hgnn_stream_t stream);
int hgnn_dense_fill_csr(const float* D1, const float* D2, long long sb, long long sr, long long sc,
                        int bs, const int* row_off, const int* col_off, const int* rowptr,
                        int* col, float* val1, float* val2, hgnn_stream_t stream);
/* out[i] = sum_{k<i} in[k], i = 0..n (n+1 outputs). */
int hgnn_exclusive_scan_i32(const int* in, int* out, int n, hgnn_stream_t stream);
/* CSR -> dense scatter into a pre-zeroed strided dense tensor (bit-exact operator checks). */
int hgnn_csr_to_dense(const int* rowptr, const int* col, const float* val, int bs,
                      const int* row_off, const int* col_off, float* D, long long sb, long long sr,
                      long long sc, hgnn_stream_t stream);
/* Block-diagonal fix-up after the raw H2D copy of a batch (hgnn-2_b200/sparse_ops.py
 * concat_block_diagonal(defer_offsets=True)): table = n_entries x (array offset, length,
 * segment-pointer offset, segment-addend offset) in 4-byte words from `base`;
 * arr[i] += addend[g] for i in [segptr[g], segptr[g+1]), g < n_seg. */
int hgnn_fixup_offsets(int* base, const int* table, int n_entries, int n_seg, hgnn_stream_t stream);
/* out[r] = sum of row r's values (weighted degree, functions/operators.py:22,74). */
int hgnn_csr_row_sums(const int* rowptr, const float* val, int R, float* out, hgnn_stream_t stream);

/* ---- SpGEMM for the power operators C = A*A (functions/operators.py:26-29, 78-81) ---------- */
/* step 1: prodcnt[r] = number of partial products of row r (then scan -> prodptr). */
int hgnn_spgemm_count_products(int R, const int* a_rowptr, const int* a_col, const int* b_rowptr,
                               int* prodcnt, hgnn_stream_t stream);
/* step 2: expand products into scratch (pcol, pval), flag first occurrences (pflag), rowcnt[r] = #unique. */
int hgnn_spgemm_expand(int R, const int* a_rowptr, const int* a_col, const float* a_val,
                       const int* b_rowptr, const int* b_col, const float* b_val,
                       const int* prodptr, int* pcol, float* pval, int* pflag, int* rowcnt,
                       hgnn_stream_t stream);
/* step 3: merge equal columns, emit CSR sorted by column; clip!=0 binarises (min(v,1)); the
 * reference does NOT clip, so the host layer passes clip=0 by default. */
int hgnn_spgemm_fill(int R, const int* prodptr, const int* pcol, const float* pval,
                     const int* pflag, const int* c_rowptr, int* c_col, float* c_val, int clip,
                     hgnn_stream_t stream);

/* ---- gmul primitives (models/layers/layers_mnb.py:391-434, functions/utils.py:24-81) ------- */
/* Y[r, t*F + f] = (ops[t] X)[r, f], t < n_ops.   `ops` is a HOST array. X: (R_in, F); Y: (R, n_ops*F).
 * graph_oper.forward: ops = [IDENT, DIAG(deg), CSR(A), CSR(A^2) ...]; P_multi.forward: one CSR. */
int hgnn_gmul_fwd(const hgnn_op_t* ops, int n_ops, int R, int F, const float* X, float* Y,
                  hgnn_stream_t stream);
/* gX[r, f] = sum_t (opsT[t] G_t)[r, f] with G_t = G[:, t*F:(t+1)*F]; opsT = transposed operators. */
int hgnn_gmul_bwd(const hgnn_op_t* opsT, int n_ops, int R, int F, const float* G, float* gX,
                  hgnn_stream_t stream);

/* ---- masked batch-norm (models/layers/batch_normalization.py:23-108) ---------------------- */
/* Per-feature batch statistics of packed rows (all R rows are real slots): stats = [mean(F),
 * std(F) = sqrt(var_biased + 1e-5), scale(F) = weight/std, shift(F) = bias - weight*mean/std];
 * running_* updated in place as 0.9*batch + 0.1*running (NULL = skip).  ws: hgnn_workspace_bytes(2F). */
int hgnn_bn_stats(const float* Z, int R, int F, const float* weight, const float* bias,
                  float* running_mean, float* running_std, float momentum, float* stats,
                  void* ws, long long ws_bytes, hgnn_stream_t stream);
/* eval mode: stats from running_mean / running_std (batch_normalization.py:39-41). */
int hgnn_bn_stats_eval(int F, const float* weight, const float* bias, const float* running_mean,
                       const float* running_std, float* stats, hgnn_stream_t stream);
/* Y = scale*Z + shift. */
int hgnn_bn_apply(const float* Z, int R, int F, const float* stats, float* Y, hgnn_stream_t stream);
/* backward, step 1: coef(3F+2): gZ = c0[f]*g + c1[f] + c2[f]*Z ; coef[3F] = d weight, coef[3F+1] =
 * d bias (scalars).  train!=0 uses batch statistics; eval mode gives c0 = scale, c1 = c2 = 0.
 * gshift (F, or NULL): gradient that reached stats.shift, i.e. the value the reference's BN leaves
 * in padded slots (batch_normalization.py:75); it is folded into the same coefficients. */
int hgnn_bn_bwd_reduce(const float* gY, const float* Z, int R, int F, const float* stats,
                       const float* weight, int train, const float* gshift, float* coef, void* ws,
                       long long ws_bytes, hgnn_stream_t stream);
/* backward, step 2 (+ ReLU mask of the conv branch): gPre[r,o] = (c0*g + c1 + c2*Z) * (o < relu_from
 * || Z > 0); coef == NULL means "no batch-norm" (gPre = g * mask).  dbias[o] = sum_r gPre[r,o].
 * ws: hgnn_workspace_bytes(2F). */
int hgnn_side_bwd_pre(const float* gY, const float* Z, int R, int F, const float* coef,
                      int relu_from, float* gPre, float* dbias, void* ws, long long ws_bytes,
                      hgnn_stream_t stream);

/* ---- fused layer side: gather -> concat -> two 1x1 convs -> ReLU -> BN statistics ----------- */
/* One "side" of a GNN / LGNN layer (models/layers/layers_mnb.py:52-69, 189-225, 256-290, 322-358,
 * readouts :88-95, :379-388):
 *   x1[r] = [ ops[0] Xs, ..., ops[n_ops-1] Xs,  Pm Xc,  Pd Xc ][r]       (Cin = n_ops*Fs + 2*Fc)
 *   Z[r]  = [ Wa x1 + ba  (Ha outputs) ,  Wb x1 + bb  (Hb outputs) ],  ReLU on outputs >= relu_from
 * The cross part is skipped when p_rowptr == NULL (then Fc must be 0).  Wa: (Ha, Cin), Wb: (Hb, Cin)
 * row-major = Conv1d(k=1).weight.  If stats != NULL the per-feature batch statistics of Z are
 * reduced in the same launch (as hgnn_bn_stats).  ws: hgnn_workspace_bytes(2*(Ha+Hb)). */
typedef struct hgnn_side_t {
    int R;                  /* output rows */
    const hgnn_op_t* ops;   /* host array of n_ops self operators */
    int n_ops;
    const float* Xs;        /* (R, Fs) */
    int Fs;
    const int* p_rowptr;    /* (R+1,) incidence rows of this side, or NULL */
    const int* p_col;
    const float* p_pm;
    const float* p_pd;
    const float* Xc;        /* (Rc, Fc) */
    int Fc;
    long long p_nnz;        /* entries of the incidence pattern, 0 = unknown (scheduling hint) */
    /* optional (hgnn_lg_side_fwd on the thread-per-row kernels only): weight of every output row in sums over rows
     * (the batch-norm statistics); rows with weight <= 0 are not computed at all.  Used for the collapsed line graph
     * (hgnn_batch_t.ew): one representative per block of identical phantom rows.  NULL = every row, weight 1. */
    const float* roww;
    const int* rowmap;      /* optional, with roww: the R rows to compute (hgnn_batch_t.erow); NULL: rows 0..R-1 */
} hgnn_side_t;

int hgnn_side_fwd(const hgnn_side_t* side, const float* Wa, const float* ba, int Ha,
                  const float* Wb, const float* bb, int Hb, int relu_from, float* Z,
                  const float* bn_weight, const float* bn_bias, float* running_mean,
                  float* running_std, float momentum, float* stats, void* ws, long long ws_bytes,
                  hgnn_stream_t stream);

/* Backward gather of one side.  With G = gPre (R_g, Fg) from hgnn_side_bwd_pre and the TRANSPOSED
 * operators (rows = the rows of X being differentiated):
 *   T[u]  = [ opsT[0] G, ..., opsT[n-1] G ][u]                              (n*Fg values)
 *   gX[u, f] (+)= sum_{t,o} W[o, col0 + t*Fx + f] * T[u, t*Fg + o]          (W = [Wa; Wb])
 *   dW[o, col0 + t*Fx + f] = sum_u T[u, t*Fg + o] * X[u, f]
 * Self part: opsT = transposed self operators, col0 = 0, X = Xs.  Cross part: two CSR ops (Pm^T,
 * Pd^T pattern of the other side), col0 = n_ops*Fs, X = Xc.  accumulate != 0 adds into gX.
 * gX may be NULL (input needs no gradient).  dWa/dWb: (Ha, Cin)/(Hb, Cin); only the columns
 * [col0, col0 + n*Fx) are written.  ws: hgnn_workspace_bytes(n*Fg*Fx). */
int hgnn_side_bwd_gather(const hgnn_op_t* opsT, int n_ops, int R, const float* G, int Fg,
                         const float* X, int Fx, const float* Wa, int Ha, const float* Wb, int Hb,
                         int Cin, int col0, float* gX, int accumulate, float* dWa, float* dWb,
                         void* ws, long long ws_bytes, hgnn_stream_t stream);

/* ---- model-level training engine (csrc/engine.cu): raw activations + on-load normalisation ---- */
/* How a consumer normalises a stored RAW (pre-batch-norm) tensor on load: either from the binned
 * fp64 (sum z, sum z^2) accumulators of its producer (training), from a precomputed
 * [scale(F), shift(F)] vector (eval), or not at all (all NULL: layer-0 inputs). */
typedef struct hgnn_bn_ref_t {
    const double* acc;
    const float* affine;
    const float* weight; /* scalar BN affine (batch_normalization.py:26-27) */
    const float* bias;
    int n_rows;
} hgnn_bn_ref_t;

/* Forward of one layer side like hgnn_side_fwd, but inputs are normalised on load (bn_self /
 * bn_cross, NULL = identity), Z is written RAW and its (sum z, sum z^2) are added to the binned
 * accumulators acc_out (hgnn_ws_bins(2*Fout) x 2*Fout doubles, zeroed by the caller; NULL = no BN).
 * No ticket, no finalisation: the consumers derive mean/std themselves.  X1 != NULL (width-4 fast
 * path only) additionally saves the concatenated, normalised input rows for hgnn_lg_side_dw. */
int hgnn_lg_side_fwd(const hgnn_side_t* side, const hgnn_bn_ref_t* bn_self, const hgnn_bn_ref_t* bn_cross,
                     const float* Wa, const float* ba, int Ha, const float* Wb, const float* bb, int Hb,
                     int relu_from, float* Z, double* acc_out, float* X1, hgnn_stream_t stream);
/* 1 if the (ops, widths) combination runs on the thread-per-row width-4 kernels (engine_row4.cuh);
 * only then may X1 (the saved concatenated input rows, (R, Cin)) and skip_dw be used. */
int hgnn_lg_row4_eligible(const hgnn_op_t* ops, int n_ops, int Fs, int Fc, int Fout);
/* 1 if a side with n_ops operators, self / cross input widths Fs / Fc (Fc = 0: no cross part) and Fout outputs
 * runs on the tensor-core tile kernels for wide states (engine_wide.cuh) - forward (backward = 0) or backward
 * (backward = 1; Fout is then the width of the incoming gradient) - given 16-byte aligned tensors.  Widths only:
 * host code and tests use it to know which code path a model takes (replaces nothing in the reference, whose
 * layers_mnb.py:189-225 has a single dense path). */
int hgnn_lg_wide_eligible(int n_ops, int Fs, int Fc, int Fout, int backward);
/* 1 when BOTH directions of a side with these widths fit the engine kernels' shared-memory budgets
 * (the weight block Cin x Fout stays resident; Cin = n_ops*Fs + 2*Fc).  Width-only; the host layer
 * (engine.supported) uses it to route too-wide models to the per-layer kernels
 * (reference: models/layers/layers_mnb.py:172-177 fixes Cin, Fout from the feature maps). */
int hgnn_lg_side_fits(int n_ops, int Fs, int Fc, int Fout);
/* Weight gradients of a width-4 side as a streaming pass over the saved x1 rows:
 * dW[o][c] += sum_r gPre[r][o] x1[r][c], dbias[o] += sum_r gPre[r][o] (gPre as in hgnn_lg_side_bwd).
 * Independent of hgnn_lg_side_bwd(skip_dw=1): the host layer runs it on a parallel stream. */
int hgnn_lg_side_dw(const float* gY, const float* Z, int R, int relu_from, const double* acc_f,
                    const double* acc_b, const float* bn_weight, const float* X1, int Cin,
                    double* dW_bins, double* db_bins, hgnn_stream_t stream);

/* Backward of one layer side in ONE launch.  gY = gradient w.r.t. the NORMALISED output of the side
 * (R_g x Fg, complete), Z its raw output.  gPre = (c0 gY + c1 + c2 Z) * relu_mask is evaluated on the
 * fly from acc_f / acc_b (acc_b == NULL: no batch-norm, gPre = gY * relu_mask).  Self part (R_self
 * rows, transposed operators ops_T, raw input Xs normalised by bn_self) and cross part (R_cross rows,
 * Pm^T/Pd^T pattern pt_*, raw input Xc): gX (+)= W^T T, dW/dbias partials to binned fp64
 * accumulators, and (sum g, sum g*xhat) of the produced gradients to acc_b_self / acc_b_cross. */
typedef struct hgnn_side_bwd_t {
    const float* gY; const float* Z; int Fg; int relu_from; int Rg;
    const double* acc_f; const double* acc_b; const float* bn_weight;
    const float* Wa; int Ha; const float* Wb; int Hb; int Cin;
    double* dW_bins; double* db_bins;
    /* self part */
    int R_self; const hgnn_op_t* ops_T; int n_ops; const float* Xs; int Fs; hgnn_bn_ref_t bn_self;
    float* gXs; int accumulate_self; double* acc_b_self;
    /* cross part (R_cross = 0: absent) */
    int R_cross; const int* pt_rowptr; const int* pt_col; const float* pt_pm; const float* pt_pd;
    const float* Xc; int Fc; hgnn_bn_ref_t bn_cross; float* gXc; int accumulate_cross; double* acc_b_cross;
    int skip_dw; /* 1: leave dW / dbias to hgnn_lg_side_dw (width-4 fast path only) */
    long long pt_nnz; /* entries of the pt_* pattern, 0 = unknown (scheduling hint) */
    /* optional: ZEROED scratch of hgnn_lg_rng_scratch_bytes(ops_T[2].rng_n) bytes.  With it the width-4 backward
     * dedicates a few CTAs of the launch to the range sums of the run-length part (they publish each sum + a ready
     * flag here) and the rows that own range entries just read them - instead of every CTA with such a row summing
     * the range itself after its row loop, a 7 us tail on 31 of 342 CTAs (profiles/logs/cta_times_phases.log). */
    void* rng_scratch;
    /* optional row weights of the self / cross rows (see hgnn_side_t.roww): rows with weight <= 0 are skipped, the
     * others enter dW, dbias and the batch-norm sums weight times.  Thread-per-row kernels only; NULL = weight 1. */
    const float* roww_self; const float* roww_cross;
    const int* rowmap_self; const int* rowmap_cross;   /* optional row lists; R_self / R_cross then count their entries */
} hgnn_side_bwd_t;
long long hgnn_lg_rng_scratch_bytes(int rng_n);
int hgnn_lg_side_bwd(const hgnn_side_bwd_t* desc, hgnn_stream_t stream);
/* profiling aid (HGNN_B200_ABLATE bit 8): (start ns, end ns, is_self) of the first n <= 2048 CTAs of the last
 * width-4 backward launch, copied to the HOST array out (3 n values); synchronises the device */
int hgnn_debug_cta_times(unsigned long long* out, int n);
/* same launch, row CTAs: (end of the row loop ns, end of the range phase ns [cross CTAs: = row loop], batch-norm /
 * gPre coefficient vectors ready ns) */
int hgnn_debug_cta_phases(unsigned long long* out, int n);
/* profiling aid (HGNN_B200_ABLATE bit 16): timeline of the width-4 side launches of a step, one slot per launch in issue
 * order, 4 values each (min CTA start, min / max "producer wait passed", max CTA end; ns).  out != NULL: the first n <= 1024
 * slots to the HOST array; reset 1: clear the slots and restart the numbering, 2: clear the slots only (a captured graph
 * keeps its slots across replays); synchronises the device */
int hgnn_debug_ktrace(unsigned long long* out, int n, int reset);

/* out[i] = sum_{b<nb[i]} sum_{c<cnt[i]} arena[off[i] + b*stride[i] + c]: every binned accumulator of
 * a step -> the flat fp32 gradient buffer, one launch. */
int hgnn_bins_reduce(const double* arena, const long long* off, const int* nb, const int* stride,
                     const int* cnt, int n, float* out, hgnn_stream_t stream);
/* Running statistics of all BN instances of a model in one launch (batch_normalization.py:37-38). */
int hgnn_bn_running_update(const double* arena, const long long* acc_off, const int* F, const int* n_rows,
                           const long long* run_off, int n_bn, float momentum, float* running,
                           hgnn_stream_t stream);
/* Readout backward prologue: G[r,o] = g[graph(r),o]; adds sum_b pad_count[b]*g[b,o] (the padded slots'
 * share of d fc.bias, layers_mnb.py:92) to bin 0 of db_bins. */
int hgnn_readout_bwd_prep(const float* g, int bs, int F, const int* off, const float* pad_count, float* G,
                          double* db_bins, hgnn_stream_t stream);
/* Number of accumulator bins used for a reduction of `width` values. */
int hgnn_bins_for(int width);

/* ---- readout (layers_mnb.py:92, :386): y[b,o] = sum_{rows of graph b} Y[r,o] + pad[b]*bias[o] */
int hgnn_segment_sum(const float* Y, int bs, int F, const int* off, const float* pad_count,
                     const float* bias, float* out, hgnn_stream_t stream);
/* G[r, o] = g[graph(r), o]  (backward of the sum). */
int hgnn_segment_bcast(const float* g, int bs, int F, const int* off, float* G,
                       hgnn_stream_t stream);

/* ---- CCN second-order covariant contraction (functions/contraction.py, utils_ccn.py) ------ */
/* Stand-alone collapse6to3 (contraction.py:106-121) on a general rank-6 tensor:
 * F6 (C, n, n, n, n, n) -> out (n, n, 18*C); and its backward gout (n, n, 18C) -> gF6. */
int hgnn_ccn2_collapse6to3(const float* F6, int C, int n, float* out, hgnn_stream_t stream);
int hgnn_ccn2_collapse6to3_bwd(const float* gout, int C, int n, float* gF6, hgnn_stream_t stream);
/* Batched per-vertex level update (utils_ccn.py:281-300 with python_contract :37-45):
 *   vertices v = 0..V-1 of all graphs; nbr_ptr (V+1), nbr (sorted receptive field, global vertex
 *   ids, includes v); nmax = largest receptive field; f_off (V+1) = prefix sums of d_v^2:
 *   F[v] is (d_v, d_v, C) at Fprev + f_off[v]*C.
 *   Fnext[v] = relu( Linear_{18C->H}( contract18( promote_{j in nbr(v)} Fprev[j],  adj = I ) ) ). */
int hgnn_ccn2_update_fwd(int V, int nmax, const int* nbr_ptr, const int* nbr, const long long* f_off,
                         const float* Fprev, int C, const float* W, const float* b, int H,
                         float* Fnext, hgnn_stream_t stream);
/* Backward: gPre = gFnext * (Fnext > 0) is formed on the fly.  gFprev (same layout as Fprev; NULL =
 * skip) is written, not accumulated; dW (H,18C) and db (H) are written.
 * ws: hgnn_workspace_bytes(H*18*C+H). */
int hgnn_ccn2_update_bwd(int V, int nmax, const int* nbr_ptr, const int* nbr, const long long* f_off,
                         const float* Fprev, int C, const float* W, int H, const float* Fnext,
                         const float* gFnext, float* gFprev, float* dW, float* db, void* ws,
                         long long ws_bytes, hgnn_stream_t stream);
/* 1-D variant (utils_ccn.py:303-324): F[v] is (d_v, C) at row nbr_ptr[v]; 2 contractions (row /
 * column sums of the promoted stack).  ws: hgnn_workspace_bytes(H*2*C+H). */
int hgnn_ccn1_update_fwd(int V, int nmax, const int* nbr_ptr, const int* nbr, const float* Fprev,
                         int C, const float* W, const float* b, int H, float* Fnext,
                         hgnn_stream_t stream);
int hgnn_ccn1_update_bwd(int V, int nmax, const int* nbr_ptr, const int* nbr, const float* Fprev,
                         int C, const float* W, int H, const float* Fnext, const float* gFnext,
                         float* gFprev, float* dW, float* db, void* ws, long long ws_bytes,
                         hgnn_stream_t stream);

/* ---- optimizer: fused Adamax step over one flat parameter buffer (scripts/main_gnn.py:160-167
 * uses torch.optim.Adamax(lr); same update rule, one launch for all parameters) -------------- */
int hgnn_adamax_step(float* param, const float* grad, float* exp_avg, float* exp_inf, long long n,
                     float lr, float beta1, float beta2, float eps, float grad_scale,
                     int* step /* device counter, incremented by the call */, hgnn_stream_t stream);

/* ---- gradient all-reduce fused with the Adamax update over NVLink peer memory (csrc/p2p.cu) ----------------
 * The reference is single-process; data-parallel training over the GPUs of one box is this repo's addition
 *  and its one exchange is the sum of the flat gradient.  Every rank allocates one peer buffer
 * (hgnn_p2p_alloc: device pointer + a 64-byte CUDA IPC handle), opens the handles of the other ranks
 * (hgnn_p2p_open) and then replaces "all-reduce, then hgnn_adamax_step" by ONE launch per step:
 * publish the local gradient, wait for the peers' flags, read their gradients over NVLink, sum in rank order,
 * apply the update.  peer_bufs: host array of `world` device pointers indexed by rank (own buffer at [rank]).
 * n <= cap_floats <= hgnn_p2p_max_floats().  *fault (device int, zeroed) is set if a peer never arrives. */
long long hgnn_p2p_buffer_bytes(long long cap_floats);
int hgnn_p2p_max_floats(void);
int hgnn_p2p_alloc(long long cap_floats, void** dev_ptr, void* handle64);
int hgnn_p2p_open(const void* handle64, void** dev_ptr);
int hgnn_p2p_close(void* dev_ptr);
int hgnn_p2p_free(void* dev_ptr);
int hgnn_p2p_allreduce_adamax(float* param, const float* grad, float* exp_avg, float* exp_inf, int n, float lr,
                              float beta1, float beta2, float eps, float grad_scale, int* step,
                              void* const* peer_bufs, int rank, int world, long long cap_floats, int* fault,
                              hgnn_stream_t stream);

/* ---- whole-model training step in two calls (csrc/program.cu) ---------------------------------
 * The layer stack of GNN_simple / GNN_lg (models/gnns/model_mnb.py:58-66, 124-129) as a static
 * "program": a table of tensors (0 = X, 1 = XL for the line-graph model, then one per layer side
 * output) and a list of sides in forward order.  hgnn_program_fwd / _bwd walk that list on the host
 * and issue the same launches as a per-side loop over hgnn_lg_side_fwd / hgnn_lg_side_bwd would,
 * without ~90 foreign-function calls per step.  Parameters are addressed by INDEX into `param_addr`,
 * a host array of device addresses in model.parameters() order. */
typedef struct hgnn_prog_tensor_t {
    int F;                  /* feature width */
    int rows;               /* 0: node rows (Rn), 1: line-graph rows (Rm) */
    int bn_weight, bn_bias; /* parameter indices of the scalar BN affine; -1: not normalised (inputs) */
    long long acc_f, acc_b; /* arena offsets (doubles) of the (sum z, sum z^2) / (sum g, sum g xhat) bins */
} hgnn_prog_tensor_t;

typedef struct hgnn_prog_side_t {
    int kind;               /* 0: node side, 1: edge (line-graph) side */
    int src_self, src_cross /* -1: none */, out /* -1: readout */;
    int Wa, ba, Ha, Wb, bb, Hb; /* parameter indices; Wb = bb = -1 when Hb == 0 */
    int relu_from;
    long long dW_off, db_off;   /* arena offsets (doubles) of the dW / dbias bins */
} hgnn_prog_side_t;

typedef struct hgnn_program_t {
    int n_tensors; const hgnn_prog_tensor_t* tensors;
    int n_sides; const hgnn_prog_side_t* sides;
    int dual;
    long long arena_doubles;
    /* device tables: accumulators -> flat gradient (hgnn_bins_reduce) */
    int n_flat; const long long* red_off; const int* red_nb; const int* red_stride; const int* red_cnt;
    /* device tables: running statistics (hgnn_bn_running_update); bn_rows_kind[k] = 0 node / 1 line-graph rows */
    int n_bn; const long long* bn_acc_off; const int* bn_F; const int* bn_rows_kind; const long long* bn_run_off;
    float momentum;
} hgnn_program_t;

typedef struct hgnn_batch_t {
    int bs, Rn, Rm, n_ops;
    const hgnn_op_t* node_ops; const hgnn_op_t* node_ops_T;
    const hgnn_op_t* edge_ops; const hgnn_op_t* edge_ops_T;          /* NULL when not dual */
    const int* p_rowptr; const int* p_col; const float* p_pm; const float* p_pd;      /* rows = nodes */
    const int* pt_rowptr; const int* pt_col; const float* pt_pm; const float* pt_pd;  /* rows = line-graph nodes */
    const int* node_off; const float* pad_n;
    long long p_nnz;      /* entries of the incidence pattern (p and pt hold the same entries) */
    /* Persistent-kernel path (csrc/mega.cu); all NULL / 0: not available, every side gets its own launch.
     * Collapsed line graph (sparse_ops.GraphOps._build_collapsed): the reference's phantom line-graph rows
     * (functions/operators.py:59,68-71) are identical copies, so one representative per graph is computed with
     * its multiplicity.  erow (n_act,) = the active rows; ew (Rm,) = row weight in sums over rows (1, the
     * multiplicity, -(distance to the representative) for skipped copies); btc_* = AL^T over the active rows with
     * the multiplicity folded in. */
    const int* btc_rowptr; const int* btc_col; const float* btc_val;
    const int* erow; const float* ew; int n_act;
    long long btc_nnz;    /* entries of btc (scheduling hint) */
    int collapse_ok;      /* 1: the edge feature XL of this call is the line-graph degree built by prepare_batch
                           * (functions/batching.py:171), i.e. identical on the copies of a phantom block */
    void* mega_scratch;   /* >= 256 bytes, zeroed once, private to the stream: grid-barrier state */
} hgnn_batch_t;

/* floats of activation workspace for a batch with Rn node rows and Rm line-graph rows (every side
 * output + the readout rows); the gradient workspace of hgnn_program_bwd has the same size. */
long long hgnn_program_work_floats(const hgnn_program_t* prog, int Rn, int Rm);
/* Training forward: zeroes `arena` (prog->arena_doubles doubles), runs every side (raw activations into
 * `work`), the readout sum into out (bs, Fout_readout) and the running-statistics update (running != NULL). */
int hgnn_program_fwd(const hgnn_program_t* prog, const hgnn_batch_t* batch, const float* X, const float* XL,
                     const long long* param_addr, float* work, double* arena, float* running, float* out,
                     hgnn_stream_t stream);
/* Backward of the same step: g_out (bs, Fout_readout) -> gX (Rn, F_X; NULL = not needed) and the flat
 * parameter gradient gflat (prog->n_flat).  gwork: scratch of hgnn_program_work_floats floats. */
int hgnn_program_bwd(const hgnn_program_t* prog, const hgnn_batch_t* batch, const float* X, const float* XL,
                     const long long* param_addr, const float* work, float* gwork, double* arena,
                     const float* g_out, float* gX, float* gflat, void* rng_scratch, long long rng_scratch_bytes,
                     hgnn_stream_t stream);
/* 1 when hgnn_program_fwd / _bwd run this (program, batch) on the collapsed line graph (batch->collapse_ok set,
 * every side on the thread-per-row kernels): the line-graph sides then compute only the active rows batch->erow,
 * weighted by batch->ew, and the backward gathers through batch->btc_*. */
int hgnn_program_uses_collapse(const hgnn_program_t* prog, const hgnn_batch_t* batch);
/* Profiling aids of the persistent kernels (csrc/mega.cu; no reference counterpart).  hgnn_mega_set_trace: device
 * buffer of 2 * 48 * grid * 4 uint64 that receives %globaltimer stamps per (phase, CTA) - phase = side index for
 * the forward, 48 + side index for the backward; stamps: phase entered, barrier passed, batch-norm vectors ready,
 * rows + flush done; NULL switches it off.  hgnn_mega_grid_for: CTAs a launch over Rn node rows and n_act active
 * line-graph rows uses. */
int hgnn_mega_set_trace(void* dev_ptr);
int hgnn_mega_grid_for(int Rn, int n_act);
/* bytes of rng_scratch for hgnn_program_bwd (0: none needed); the call zeroes it itself; NULL / too small: without */
long long hgnn_program_rng_scratch_bytes(const hgnn_program_t* prog, const hgnn_batch_t* batch);
/* kernels launched by hgnn_program_* calls so far (the host layer adds it to its own launch count) */
long long hgnn_program_launches(void);
/* hgnn_bn_running_update with the row counts given as (kind table, Rn, Rm) instead of a device vector */
int hgnn_bn_running_update_k(const double* arena, const long long* acc_off, const int* F, const int* rows_kind,
                             int Rn, int Rm, const long long* run_off, int n_bn, float momentum, float* running,
                             hgnn_stream_t stream);

/* ---- host side of the batch pack (csrc/hostpack.cu; plain CPU code, no CUDA calls) ---------------
 * functions/batching.py:77-185 zero-pads and stacks dense per-graph operators; here every graph carries
 * ONE contiguous host blob of its CSR arrays (sparse_ops.GraphOps.blob) and a batch is their
 * block-diagonal concatenation into one (pinned) staging buffer, laid out exactly like
 * sparse_ops.concat_block_diagonal(defer_offsets=True): raw per-field concatenation + segment tables
 * + the fix-up table that hgnn_fixup_offsets applies on the GPU after the copy.
 * Blob = int64 header [HGNN_BLOB_MAGIC, N, M, E, n_fields, (byte offset, length in 4-byte elements) x n_fields]
 * followed by the arrays, each 16-byte aligned, in the field order of HGNN_BLOB_FIELDS (hostpack.cu). */
#define HGNN_BLOB_MAGIC 0x48474e4e424c4f42ll
/* Wall-clock split of the last hgnn_host_pack_fill call in ns: which = 0 task list, 1 copies (profiling aid). */
long long hgnn_host_pack_last_ns(int which);
/* Number of output arrays of a batch layout (`layout` below holds that many (offset, length) pairs). */
int hgnn_host_pack_n_keys(void);
/* Name of output array k (static string; the Python layer builds its view table from these). */
const char* hgnn_host_pack_key(int k);
/* Computes the batch layout: layout[2k] = byte offset, layout[2k+1] = length (4-byte elements) of array k
 * (length -1: array absent, e.g. the line-graph arrays when dual == 0 or bt when skip_bt != 0).
 * Returns the total number of bytes (16-byte aligned sub-arrays), < 0 on error. */
long long hgnn_host_pack_layout(int bs, const void* const* blobs, int dual, int skip_bt, long long* layout);
/* Padded host feature tensors of prepare_batch (functions/batching.py:113-127,:171 of the reference): X (bs, n_feat, Nmax)
 * with X[g, f, j] = x_rows[g][j * n_feat + f] for j < N_g and zero beyond; XL (bs, 1, Emax) = the line-graph degree of every
 * graph, read from its blob (NULL: primal batch).  Host code only. */
int hgnn_host_fill_features(int bs, const void* const* blobs, const float* const* x_rows, int n_feat, long long Nmax,
                            float* X, long long Emax, float* XL);

/* Fills `out` (total bytes from hgnn_host_pack_layout) using up to n_threads host threads. */
int hgnn_host_pack_fill(int bs, const void* const* blobs, int dual, int skip_bt, const long long* layout,
                        void* out, int n_threads);

/* ---- device-side batch assembly (the product path of prepare_batch) ------------------------------
 * Every graph blob goes host->device as it is (one cudaMemcpyAsync per graph, straight from the dataset's
 * pinned memory), then one kernel gathers the fields into the block-diagonal arrays and adds the per-graph
 * row / column / nnz offsets on the way: same arrays as hgnn_host_pack_fill + hgnn_fixup_offsets, without
 * the host-side concatenation (functions/batching.py:77-185 is the reference's dense zero-padding).
 * hgnn_pack_device_plan: layout[2k], layout[2k+1] = byte offset / length of output array k (keys of
 * hgnn_host_pack_key; segment tables and the fix-up table are absent: length -1); returns the bytes of the
 * output buffer, *stage_bytes = device staging for the raw blobs, *meta_bytes = size of the task table
 * (needed twice: pinned host + device).  < 0 on error. */
long long hgnn_pack_device_plan(int bs, const void* const* blobs, int dual, int skip_bt, long long* layout,
                                long long* stage_bytes, long long* meta_bytes);
/* Enqueues the copies and the gather kernel on `stream`.  meta_host must stay untouched, and the blobs alive,
 * until the stream has passed this point. */
int hgnn_pack_device_upload(int bs, const void* const* blobs, int dual, int skip_bt, void* out_dev,
                            void* stage_dev, void* meta_host, void* meta_dev, hgnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HGNN_B200_H */
