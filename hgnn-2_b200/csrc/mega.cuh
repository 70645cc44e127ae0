// mega.cuh -- parameter blocks of the persistent ("mega") engine kernels in mega.cu, shared with the step
// executor (program.cu), which fills them.
//
// One cooperative kernel per pass runs a whole run of width-4 layer sides of a GNN_simple / GNN_lg model
// (models/gnns/model_mnb.py:58-66,124-129 over layers_mnb.py:52-69,189-225): every CTA owns a fixed slice of
// the node rows and of the ACTIVE line-graph rows for all layers, sides are separated by a grid barrier (the
// batch-norm statistics of a side must be complete before its consumers normalise on load), activations stay
// in L2, the graph structure of a CTA's slice stays in its L1.
#pragma once
#include <cuda_runtime.h>

namespace mk {

constexpr int MAX_SIDES = 48;
constexpr int MAX_TENSORS = 50;
constexpr int THREADS = 512;

struct Tensor {
    float* data;            // raw (pre-batch-norm) values, (rows, 4)
    float* grad;            // backward: gradient w.r.t. the NORMALISED tensor, (rows, 4)
    const double* acc_f;    // binned (sum z, sum z^2) of the producer; NULL: not normalised
    double* acc_b;          // binned (sum g, sum g xhat); NULL: not normalised
    const float* bn_w;      // scalar batch-norm affine
    const float* bn_b;
    int n_rows;             // rows behind the statistics (ALL rows, phantom copies included)
    int pad;
    double inv_n;           // 1 / n_rows
};

struct Side {
    int kind;               // 0: node rows, 1: line-graph rows
    int src_self, src_cross /* -1: none */, out;
    const float* Wa; const float* ba; const float* Wb; const float* bb;
    int Ha, Hb, relu_from, Cin;
    double* dW_bins; double* db_bins;
    int need_self, need_cross, acc_self, acc_cross;     // backward only
};

// Block-diagonal structure of the batch.  Line-graph rows are "collapsed" (sparse_ops.GraphOps._build_collapsed):
// erow lists the n_act active rows, ew[r] is the weight of row r in every sum over rows (1, the multiplicity for
// the representative of a phantom block, and -(distance to the representative) for the skipped copies).
struct Graph {
    int Rn, Rm, n_act, pad;
    const float* deg; const int* a_rp; const int* a_col; const float* a_val;      // node rows, forward
    const int* at_rp; const int* at_col; const float* at_val;                    // node rows, backward (A^T)
    const float* dl; const int* b_rp; const int* b_col; const float* b_val;       // line-graph rows, forward
    const int* btc_rp; const int* btc_col; const float* btc_val;                 // line-graph rows, backward
    const int* p_rp; const int* p_col; const float* p_pm; const float* p_pd;      // incidence, rows = nodes
    const int* pt_rp; const int* pt_col; const float* pt_pm; const float* pt_pd;  // incidence, rows = line graph
    const int* erow; const float* ew;
};

struct Params {
    Graph g;
    int n_tensors, n_sides;
    int expand;             // tensor whose skipped line-graph rows get the representative's value at the end
                            // (forward: data, backward: grad) because a per-side kernel reads it next; -1: none
    int pad;
    int grid, max_n, max_e, cap_words;   // launch plan (hgnn_mega_plan): CTAs, header slots per CTA, entry words of the cache
    long long nnz1_n, nnz1_e, nnz2;      // entries of the node / line-graph CSR operator and of the incidence pattern (hints; -1 unknown)
    unsigned int* bar;      // grid barrier state: [0] arrival counter, [32] exit counter (zero-initialised once, left zero)
    unsigned long long* trace;  // profiling aid (hgnn_mega_set_trace): 4 x %globaltimer per (phase, CTA); NULL = off
    Tensor t[MAX_TENSORS];
    Side s[MAX_SIDES];
};

}  // namespace mk

// launch wrappers (mega.cu); return HGNN_OK or an error code with hgnn_last_error set
unsigned long long* hgnn_mega_trace_ptr(void);
void hgnn_mega_plan(mk::Params* p);     // fills grid / max_n / max_e / cap_words from g and the nnz hints
int hgnn_mega_launch_fwd(const mk::Params& p, cudaStream_t stream);
int hgnn_mega_launch_bwd(const mk::Params& p, cudaStream_t stream);
