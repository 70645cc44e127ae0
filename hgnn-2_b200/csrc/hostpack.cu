// hostpack.cu -- host side of the batch pack: block-diagonal concatenation of per-graph CSR blobs
// into one (pinned) staging buffer, multi-threaded.  Plain CPU code (no CUDA calls).
//
// Reference: functions/batching.py:77-185 zero-pads dense per-graph operators into (bs, N, N, K) /
// (bs, M, M, K) tensors with torch.cat / copy_ - 9.6 GB of WL for 32 graphs of N=1000.  Here a batch
// is ~19 MB of CSR arrays; the per-field concatenation below is what remains of prepare_batch on
// the host.  The output layout equals sparse_ops.concat_block_diagonal(defer_offsets=True) array for
// array (tests/test_hostpack.py compares them bit for bit): raw per-field concatenation, segment
// tables, and the fix-up table that hgnn_fixup_offsets applies on the GPU after the copy.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <vector>
#include "common.cuh"

namespace {

// ---- blob fields (order = sparse_ops.BLOB_FIELDS) ----------------------------------------------
enum Field {
    F_DEG = 0, F_A_RP, F_A_COL, F_A_VAL, F_AT_RP, F_AT_COL, F_AT_VAL, N_PRIMAL_FIELDS,
    F_DL = N_PRIMAL_FIELDS, F_B_RP, F_B_COL, F_B_VAL,
    F_P_RP, F_P_COL, F_P_PM, F_P_PD, F_PT_RP, F_PT_COL, F_PT_PM, F_PT_PD,
    F_BTS_RP, F_BTS_COL, F_BTS_VAL, F_RNG_RP, F_RNG_ID, F_RNG_VAL, F_RNG_LO, F_RNG_HI,
    F_BTC_RP, F_BTC_COL, F_BTC_VAL, F_EW, F_EROW,     // collapsed line graph (sparse_ops.GraphOps._build_collapsed)
    F_BT_RP, F_BT_COL, F_BT_VAL,      // last: a batch that skips the full transposed operator copies a prefix
    N_FIELDS
};

struct Blob {
    const long long* h;   // header
    const char* base;
    long long N() const { return h[1]; }
    long long M() const { return h[2]; }
    int n_fields() const { return (int)h[4]; }
    long long len(int f) const { return h[5 + 2 * f + 1]; }
    const char* ptr(int f) const { return base + h[5 + 2 * f]; }
};

// ---- output arrays --------------------------------------------------------------------------------
enum KeyType { T_NODE_OFF, T_EDGE_OFF, T_PAD_N, T_SEG, T_RAW, T_RP, T_FIXUP };
enum Group { G_PRIMAL, G_DUAL, G_BT };

struct KeyDesc {
    const char* name;
    KeyType type;
    int field;     // T_RAW / T_RP: source field; T_SEG: field whose lengths are accumulated
    Group group;
};

// clang-format off
const KeyDesc KEYS[] = {
    {"node_off", T_NODE_OFF, -1, G_PRIMAL},          // 0
    {"edge_off", T_EDGE_OFF, -1, G_PRIMAL},          // 1
    {"pad_n", T_PAD_N, -1, G_PRIMAL},                // 2
    {"deg", T_RAW, F_DEG, G_PRIMAL},                 // 3
    {"dl", T_RAW, F_DL, G_DUAL},                     // 4
    {"_seg_rng_entries", T_SEG, F_RNG_ID, G_DUAL},   // 5
    {"_seg_rng_ranges", T_SEG, F_RNG_LO, G_DUAL},    // 6
    {"bts_rng_rowptr", T_RP, F_RNG_RP, G_DUAL},      // 7
    {"bts_rng_id", T_RAW, F_RNG_ID, G_DUAL},         // 8
    {"bts_rng_val", T_RAW, F_RNG_VAL, G_DUAL},       // 9
    {"bts_rng_lo", T_RAW, F_RNG_LO, G_DUAL},         // 10
    {"bts_rng_hi", T_RAW, F_RNG_HI, G_DUAL},         // 11
    {"_seg_nnz_a", T_SEG, F_A_COL, G_PRIMAL},        // 12
    {"a_rowptr", T_RP, F_A_RP, G_PRIMAL},            // 13
    {"a_col", T_RAW, F_A_COL, G_PRIMAL},             // 14
    {"a_val", T_RAW, F_A_VAL, G_PRIMAL},             // 15
    {"_seg_nnz_at", T_SEG, F_AT_COL, G_PRIMAL},      // 16
    {"at_rowptr", T_RP, F_AT_RP, G_PRIMAL},          // 17
    {"at_col", T_RAW, F_AT_COL, G_PRIMAL},           // 18
    {"at_val", T_RAW, F_AT_VAL, G_PRIMAL},           // 19
    {"_seg_nnz_b", T_SEG, F_B_COL, G_DUAL},          // 20
    {"b_rowptr", T_RP, F_B_RP, G_DUAL},              // 21
    {"b_col", T_RAW, F_B_COL, G_DUAL},               // 22
    {"b_val", T_RAW, F_B_VAL, G_DUAL},               // 23
    {"_seg_nnz_bt", T_SEG, F_BT_COL, G_BT},          // 24
    {"bt_rowptr", T_RP, F_BT_RP, G_BT},              // 25
    {"bt_col", T_RAW, F_BT_COL, G_BT},               // 26
    {"bt_val", T_RAW, F_BT_VAL, G_BT},               // 27
    {"_seg_nnz_p", T_SEG, F_P_COL, G_DUAL},          // 28
    {"p_rowptr", T_RP, F_P_RP, G_DUAL},              // 29
    {"p_col", T_RAW, F_P_COL, G_DUAL},               // 30
    {"p_pm", T_RAW, F_P_PM, G_DUAL},                 // 31
    {"p_pd", T_RAW, F_P_PD, G_DUAL},                 // 32
    {"_seg_nnz_pt", T_SEG, F_PT_COL, G_DUAL},        // 33
    {"pt_rowptr", T_RP, F_PT_RP, G_DUAL},            // 34
    {"pt_col", T_RAW, F_PT_COL, G_DUAL},             // 35
    {"pt_pm", T_RAW, F_PT_PM, G_DUAL},               // 36
    {"pt_pd", T_RAW, F_PT_PD, G_DUAL},               // 37
    {"_seg_nnz_bts", T_SEG, F_BTS_COL, G_DUAL},      // 38
    {"bts_rowptr", T_RP, F_BTS_RP, G_DUAL},          // 39
    {"bts_col", T_RAW, F_BTS_COL, G_DUAL},           // 40
    {"bts_val", T_RAW, F_BTS_VAL, G_DUAL},           // 41
    {"_seg_nnz_btc", T_SEG, F_BTC_COL, G_DUAL},      // 42
    {"btc_rowptr", T_RP, F_BTC_RP, G_DUAL},          // 43
    {"btc_col", T_RAW, F_BTC_COL, G_DUAL},           // 44
    {"btc_val", T_RAW, F_BTC_VAL, G_DUAL},           // 45
    {"ew", T_RAW, F_EW, G_DUAL},                     // 46
    {"_seg_erow", T_SEG, F_EROW, G_DUAL},            // 47
    {"erow", T_RAW, F_EROW, G_DUAL},                 // 48
    {"fixup", T_FIXUP, -1, G_PRIMAL},                // 49
};
constexpr int N_KEYS = sizeof(KEYS) / sizeof(KEYS[0]);
enum { K_NODE_OFF = 0, K_EDGE_OFF = 1, K_FIXUP = N_KEYS - 1 };

// index arrays and the (segment pointers, per-segment addend) pair that globalises them on the GPU:
// arr[i] += addend[g] for i in [segptr[g], segptr[g+1])
struct FixDesc { int arr, segp, sadd; };
const FixDesc FIXES[] = {
    {7, K_EDGE_OFF, 5}, {8, 5, 6}, {10, 6, K_EDGE_OFF}, {11, 6, K_EDGE_OFF},
    {13, K_NODE_OFF, 12}, {14, 12, K_NODE_OFF},          // a : rows n, cols n
    {17, K_NODE_OFF, 16}, {18, 16, K_NODE_OFF},          // at
    {21, K_EDGE_OFF, 20}, {22, 20, K_EDGE_OFF},          // b : rows m, cols m
    {25, K_EDGE_OFF, 24}, {26, 24, K_EDGE_OFF},          // bt
    {29, K_NODE_OFF, 28}, {30, 28, K_EDGE_OFF},          // p : rows n, cols m
    {34, K_EDGE_OFF, 33}, {35, 33, K_NODE_OFF},          // pt: rows m, cols n
    {39, K_EDGE_OFF, 38}, {40, 38, K_EDGE_OFF},          // bts
    {43, K_EDGE_OFF, 42}, {44, 42, K_EDGE_OFF},          // btc
    {48, 47, K_EDGE_OFF},                                // erow: active line-graph rows
};
// clang-format on
constexpr int N_FIXES = sizeof(FIXES) / sizeof(FIXES[0]);

inline bool present(const KeyDesc& k, int dual, int skip_bt) {
    if (k.group == G_PRIMAL) return true;
    if (!dual) return false;
    return !(k.group == G_BT && skip_bt);
}

bool open_blobs(int bs, const void* const* blobs, int dual, std::vector<Blob>* out) {
    out->resize(bs);
    for (int g = 0; g < bs; ++g) {
        const long long* h = static_cast<const long long*>(blobs[g]);
        if (!h || h[0] != HGNN_BLOB_MAGIC) return false;
        const int nf = (int)h[4];
        if (nf != N_PRIMAL_FIELDS && nf != N_FIELDS) return false;
        if (dual && nf != N_FIELDS) return false;
        (*out)[g] = Blob{h, reinterpret_cast<const char*>(h)};
    }
    return true;
}

inline long long key_length(const KeyDesc& k, const std::vector<Blob>& B, int n_fix) {
    const int bs = (int)B.size();
    long long n = 0;
    switch (k.type) {
        case T_NODE_OFF: case T_EDGE_OFF: case T_SEG: return bs + 1;
        case T_PAD_N: return bs;
        case T_RAW: for (const Blob& b : B) n += b.len(k.field); return n;
        case T_RP: for (const Blob& b : B) n += b.len(k.field) - 1; return n + 1;
        case T_FIXUP: return 4ll * n_fix;
    }
    return 0;
}

inline int count_fixes(int dual, int skip_bt) {
    int n = 0;
    for (int i = 0; i < N_FIXES; ++i) n += present(KEYS[FIXES[i].arr], dual, skip_bt) ? 1 : 0;
    return n;
}

// ---- a small persistent worker pool ---------------------------------------------------------------
class Pool {
  public:
    // runs fn(worker) for worker in [0, n); the caller is worker 0
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 1) { fn(0); return; }
        std::lock_guard<std::mutex> serial(run_mutex_);
        ensure(n - 1);
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = &fn;
            active_ = n - 1;
            pending_ = n - 1;
            ++gen_;
        }
        go_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    void ensure(int n_workers) {
        if (pid_ != getpid()) {   // forked child: the parent's threads do not exist here
            for (auto& t : threads_) t.detach();
            threads_.clear();
            pid_ = getpid();
        }
        while ((int)threads_.size() < n_workers) {
            const int idx = (int)threads_.size() + 1;
            const unsigned long long seen = gen_;
            threads_.emplace_back([this, idx, seen] { loop(idx, seen); });
        }
    }
    void loop(int idx, unsigned long long seen) {
        for (;;) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> l(m_);
                go_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (idx <= active_) job = job_;
            }
            if (job) {
                (*job)(idx);
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex run_mutex_, m_;
    std::condition_variable go_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)>* job_ = nullptr;
    unsigned long long gen_ = 0;
    int active_ = 0, pending_ = 0;
    pid_t pid_ = 0;
};

Pool& pool() {
    static Pool* p = new Pool();   // never destroyed: worker threads may outlive static destructors
    return *p;
}

struct CopyTask { char* dst; const char* src; size_t bytes; };

// wall-clock split of the last hgnn_host_pack_fill call: [0] task list, [1] copies (pool), in ns
long long g_last_ns[5] = {0, 0, 0, 0, 0};   // [2] latest worker start, [3] earliest worker start, [4] bytes copied by the caller
inline long long now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1000000000ll + ts.tv_nsec;
}

}  // namespace

extern "C" long long hgnn_host_pack_last_ns(int which) { return (which >= 0 && which < 5) ? g_last_ns[which] : -1; }

extern "C" int hgnn_host_pack_n_keys(void) { return N_KEYS; }

extern "C" const char* hgnn_host_pack_key(int k) { return (k >= 0 && k < N_KEYS) ? KEYS[k].name : nullptr; }

extern "C" long long hgnn_host_pack_layout(int bs, const void* const* blobs, int dual, int skip_bt,
                                           long long* layout) {
    if (bs < 0 || (bs > 0 && !blobs) || !layout) {
        hgnn_set_error("hgnn_host_pack_layout: bad argument");
        return HGNN_ERR_ARG;
    }
    std::vector<Blob> B;
    if (!open_blobs(bs, blobs, dual, &B)) {
        hgnn_set_error("hgnn_host_pack_layout: malformed graph blob (magic / field count)");
        return HGNN_ERR_ARG;
    }
    const int n_fix = count_fixes(dual, skip_bt);
    long long total = 0;
    for (int k = 0; k < N_KEYS; ++k) {
        if (!present(KEYS[k], dual, skip_bt)) {
            layout[2 * k] = 0;
            layout[2 * k + 1] = -1;
            continue;
        }
        const long long len = key_length(KEYS[k], B, n_fix);
        layout[2 * k] = total;
        layout[2 * k + 1] = len;
        total += (4 * len + 15) & ~15ll;
    }
    return total < 16 ? 16 : total;
}

extern "C" int hgnn_host_pack_fill(int bs, const void* const* blobs, int dual, int skip_bt, const long long* layout,
                                   void* out, int n_threads) {
    HGNN_REQUIRE(bs >= 0 && (bs == 0 || blobs) && layout && out, "bad argument");
    const long long t_start = now_ns();
    std::vector<Blob> B;
    HGNN_REQUIRE(open_blobs(bs, blobs, dual, &B), "malformed graph blob (magic / field count)");
    char* base = static_cast<char*>(out);
    auto ivec = [&](int k) { return reinterpret_cast<int*>(base + layout[2 * k]); };
    std::vector<CopyTask> tasks;
    tasks.reserve((size_t)N_KEYS * (bs + 1));
    long long nmax = 0;
    for (const Blob& b : B) nmax = b.N() > nmax ? b.N() : nmax;
    for (int k = 0; k < N_KEYS; ++k) {
        const KeyDesc& K = KEYS[k];
        if (layout[2 * k + 1] < 0) continue;
        switch (K.type) {
            case T_NODE_OFF: case T_EDGE_OFF: {
                int* o = ivec(k);
                long long c = 0;
                for (int g = 0; g < bs; ++g) { o[g] = (int)c; c += K.type == T_NODE_OFF ? B[g].N() : B[g].M(); }
                o[bs] = (int)c;
                break;
            }
            case T_PAD_N: {
                float* o = reinterpret_cast<float*>(base + layout[2 * k]);
                for (int g = 0; g < bs; ++g) o[g] = (float)(nmax - B[g].N());
                break;
            }
            case T_SEG: {
                int* o = ivec(k);
                long long c = 0;
                for (int g = 0; g < bs; ++g) { o[g] = (int)c; c += B[g].len(K.field); }
                o[bs] = (int)c;
                break;
            }
            case T_RAW: case T_RP: {
                char* dst = base + layout[2 * k];
                const int drop = K.type == T_RP ? 1 : 0;
                for (int g = 0; g < bs; ++g) {
                    const size_t bytes = 4 * (size_t)(B[g].len(K.field) - drop);
                    if (bytes) tasks.push_back(CopyTask{dst, B[g].ptr(K.field), bytes});
                    dst += bytes;
                }
                if (drop) {   // closing row pointer = total number of entries (lengths of the next field)
                    long long tot = 0;
                    for (int g = 0; g < bs; ++g) tot += B[g].len(K.field + 1);
                    *reinterpret_cast<int*>(dst) = (int)tot;
                }
                break;
            }
            case T_FIXUP: {
                int* t = ivec(k);
                for (int i = 0; i < N_FIXES; ++i) {
                    const FixDesc& f = FIXES[i];
                    if (layout[2 * f.arr + 1] < 0) continue;
                    const long long body = layout[2 * f.arr + 1] - (KEYS[f.arr].type == T_RP ? 1 : 0);
                    t[0] = (int)(layout[2 * f.arr] / 4);
                    t[1] = (int)body;
                    t[2] = (int)(layout[2 * f.segp] / 4);
                    t[3] = (int)(layout[2 * f.sadd] / 4);
                    t += 4;
                }
                break;
            }
        }
    }
    // split the copies into <= 256 KiB chunks handed out through one atomic counter
    const size_t CHUNK = 256 << 10;
    std::vector<CopyTask> chunks;
    size_t total = 0;
    for (const CopyTask& t : tasks) {
        for (size_t o = 0; o < t.bytes; o += CHUNK)
            chunks.push_back(CopyTask{t.dst + o, t.src + o, t.bytes - o < CHUNK ? t.bytes - o : CHUNK});
        total += t.bytes;
    }
    int nt = n_threads < 1 ? 1 : n_threads;
    const int by_size = (int)(total / (512 << 10)) + 1;   // no point in waking a thread for < 512 KiB
    if (nt > by_size) nt = by_size;
    std::atomic<size_t> next{0};
    const long long t_copy = now_ns();
    std::atomic<long long> first_start{1ll << 62}, last_start{0}, main_bytes{0};
    auto work = [&](int w) {
        if (w > 0) {
            const long long d = now_ns() - t_copy;
            long long cur = last_start.load();
            while (d > cur && !last_start.compare_exchange_weak(cur, d)) {}
            cur = first_start.load();
            while (d < cur && !first_start.compare_exchange_weak(cur, d)) {}
        }
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= chunks.size()) break;
            memcpy(chunks[i].dst, chunks[i].src, chunks[i].bytes);
            if (w == 0) main_bytes.fetch_add((long long)chunks[i].bytes, std::memory_order_relaxed);
        }
    };
    pool().run(nt, work);
    g_last_ns[2] = last_start.load();
    g_last_ns[3] = first_start.load() == (1ll << 62) ? 0 : first_start.load();
    g_last_ns[4] = main_bytes.load();
    g_last_ns[0] = t_copy - t_start;
    g_last_ns[1] = now_ns() - t_copy;
    return HGNN_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-side batch assembly: every graph blob is copied host->device AS IT IS (one cudaMemcpyAsync
// per graph, straight out of the dataset's pinned memory - no host-side concatenation at all), then
// ONE kernel gathers the fields into the block-diagonal arrays and adds the per-graph row / column /
// nnz offsets on the way.  The host only builds a table of (src, dst, n, addend) copy tasks.
// The result equals hgnn_host_pack_fill + hgnn_fixup_offsets bit for bit (tests/test_gpu_pack.py).
// ---------------------------------------------------------------------------------------------
namespace {

struct PackTask {
    const int* src;
    int* dst;
    int n;
    int addend;
};
static_assert(sizeof(PackTask) == 24, "PackTask layout");

__global__ void __launch_bounds__(256) pack_gather_kernel(const PackTask* __restrict__ tasks, int n_tasks) {
    for (int t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const PackTask k = tasks[t];
        int i = threadIdx.x;
        for (; i + 3 * 256 < k.n; i += 4 * 256) {   // four independent loads in flight per thread
            const int a = __ldg(k.src + i), b = __ldg(k.src + i + 256), c = __ldg(k.src + i + 512),
                      d = __ldg(k.src + i + 768);
            k.dst[i] = a + k.addend;
            k.dst[i + 256] = b + k.addend;
            k.dst[i + 512] = c + k.addend;
            k.dst[i + 768] = d + k.addend;
        }
        for (; i < k.n; i += 256) k.dst[i] = __ldg(k.src + i) + k.addend;
    }
}

inline long long blob_bytes(const Blob& b, int skip_bt = 0) {
    if (skip_bt && b.n_fields() == N_FIELDS) return b.h[5 + 2 * F_BT_RP];
    const int last = b.n_fields() - 1;
    return b.h[5 + 2 * last] + ((4 * b.len(last) + 15) & ~15ll);
}

// what is added to an index array of graph g: nothing, its node / line-graph row offset, or the number of
// entries of `field` in the graphs before it
enum AddKind { ADD_NONE = -3, ADD_NODE = -2, ADD_EDGE = -1 };
inline int add_kind(int key) {
    for (int i = 0; i < N_FIXES; ++i) {
        if (FIXES[i].arr != key) continue;
        if (FIXES[i].sadd == K_NODE_OFF) return ADD_NODE;
        if (FIXES[i].sadd == K_EDGE_OFF) return ADD_EDGE;
        return KEYS[FIXES[i].sadd].field;
    }
    return ADD_NONE;
}

inline bool on_device_path(const KeyDesc& k, int dual, int skip_bt) {
    return present(k, dual, skip_bt) && k.type != T_SEG && k.type != T_FIXUP;
}

// Offsets of the graph blobs inside the device staging buffer.  Consecutive graphs of a dataset sit back to back in the
// pinned slabs (64-byte aligned), so a batch of neighbours is ONE contiguous host range: blobs that follow their
// predecessor within HGNN_STAGE_MAX_GAP bytes keep their host spacing (the alignment gap rides along) and the whole run
// goes up in one DMA - 512 QM9-sized graphs are a handful of copies instead of 512 x ~2 us of cudaMemcpyAsync.  Anything
// else (shuffled batches, blobs whose transposed-operator tail is skipped) starts a new run.  off has bs + 1 entries:
// off[bs] = bytes of staging needed; run_start[g] = 1 where a new DMA begins.
#define HGNN_STAGE_MAX_GAP 512
long long stage_offsets(const std::vector<Blob>& B, int skip_bt, std::vector<long long>* off, std::vector<char>* run_start) {
    const int bs = (int)B.size();
    off->assign(bs + 1, 0);
    if (run_start) run_start->assign(bs + 1, 0);
    long long end = 0;      // end of the previous blob in the staging buffer
    for (int g = 0; g < bs; ++g) {
        const long long bytes = blob_bytes(B[g], skip_bt);
        bool joined = false;
        if (g > 0) {
            const long long prev_bytes = blob_bytes(B[g - 1], skip_bt);
            const long long gap = (long long)(B[g].base - B[g - 1].base) - prev_bytes;
            // both bases are 64-byte aligned in the slabs; a distance that keeps 16-byte alignment keeps the fields aligned
            if (B[g].base > B[g - 1].base && gap >= 0 && gap <= HGNN_STAGE_MAX_GAP && ((prev_bytes + gap) & 15) == 0) {
                (*off)[g] = (*off)[g - 1] + prev_bytes + gap;
                joined = true;
            }
        }
        if (!joined) {
            (*off)[g] = (end + 15) & ~15ll;
            if (run_start) (*run_start)[g] = 1;
        }
        end = (*off)[g] + bytes;
    }
    (*off)[bs] = end;
    if (run_start) (*run_start)[bs] = 1;
    return end < 16 ? 16 : end;
}

struct DevicePlan {
    long long out_bytes = 0, stage_bytes = 0, meta_bytes = 0;
    int n_tasks = 0, n_small = 0;   // n_small: ints of small arrays kept in the meta buffer
    long long small_off = 0;        // byte offset of the small arrays inside the meta buffer
    // Many small blobs that are NOT neighbours in host memory (a shuffled batch of QM9-sized graphs: 512 x ~1 KB) would be
    // hundreds of ~2 us DMAs: they are copied into the pinned meta buffer behind the task table instead (a few tens of
    // microseconds of memcpy) and ride on its ONE host->device copy.  blob_meta_off: where they start in the meta buffer.
    bool stage_in_meta = false;
    long long blob_meta_off = 0;
    std::vector<long long> blob_off;     // per blob: offset inside the staging area (device staging buffer, or meta)
    std::vector<char> run_start;         // per blob: 1 where a new DMA begins (device staging mode)
};
#define HGNN_STAGE_IN_META_MAX_BYTES (4ll << 20)
#define HGNN_STAGE_IN_META_MIN_RUNS 8

DevicePlan plan_device(const std::vector<Blob>& B, int dual, int skip_bt, long long* layout) {
    DevicePlan p;
    const int bs = (int)B.size();
    for (int k = 0; k < N_KEYS; ++k) {
        if (!on_device_path(KEYS[k], dual, skip_bt)) {
            layout[2 * k] = 0;
            layout[2 * k + 1] = -1;
            continue;
        }
        const long long len = key_length(KEYS[k], B, 0);
        layout[2 * k] = p.out_bytes;
        layout[2 * k + 1] = len;
        p.out_bytes += (4 * len + 15) & ~15ll;
        switch (KEYS[k].type) {
            case T_NODE_OFF: case T_EDGE_OFF: p.n_tasks += 1; p.n_small += bs + 1; break;
            case T_PAD_N: p.n_tasks += 1; p.n_small += bs; break;
            case T_RAW: p.n_tasks += bs; break;
            case T_RP: p.n_tasks += bs + 1; p.n_small += 1; break;
            default: break;
        }
    }
    if (p.out_bytes < 16) p.out_bytes = 16;
    p.stage_bytes = stage_offsets(B, skip_bt, &p.blob_off, &p.run_start);
    p.small_off = ((long long)p.n_tasks * sizeof(PackTask) + 15) & ~15ll;
    p.meta_bytes = p.small_off + 4ll * p.n_small + 16;
    int runs = 0;
    for (int g = 0; g < bs; ++g) runs += p.run_start[g] ? 1 : 0;
    if (runs >= HGNN_STAGE_IN_META_MIN_RUNS && p.stage_bytes <= HGNN_STAGE_IN_META_MAX_BYTES) {
        long long end = 0;                       // tight layout, 16-byte aligned blobs
        for (int g = 0; g < bs; ++g) {
            p.blob_off[g] = (end + 15) & ~15ll;
            end = p.blob_off[g] + blob_bytes(B[g], skip_bt);
        }
        p.blob_off[bs] = end;
        p.stage_in_meta = true;
        p.blob_meta_off = (p.meta_bytes + 15) & ~15ll;
        p.meta_bytes = p.blob_meta_off + end + 16;
        p.stage_bytes = 16;                      // the separate device staging buffer is not used
    }
    return p;
}

}  // namespace

extern "C" long long hgnn_pack_device_plan(int bs, const void* const* blobs, int dual, int skip_bt, long long* layout,
                                           long long* stage_bytes, long long* meta_bytes) {
    if (bs < 0 || (bs > 0 && !blobs) || !layout || !stage_bytes || !meta_bytes) {
        hgnn_set_error("hgnn_pack_device_plan: bad argument");
        return HGNN_ERR_ARG;
    }
    std::vector<Blob> B;
    if (!open_blobs(bs, blobs, dual, &B)) {
        hgnn_set_error("hgnn_pack_device_plan: malformed graph blob (magic / field count)");
        return HGNN_ERR_ARG;
    }
    const DevicePlan p = plan_device(B, dual, skip_bt, layout);
    *stage_bytes = p.stage_bytes;
    *meta_bytes = p.meta_bytes;
    return p.out_bytes;
}

extern "C" int hgnn_pack_device_upload(int bs, const void* const* blobs, int dual, int skip_bt, void* out_dev,
                                       void* stage_dev, void* meta_host, void* meta_dev, hgnn_stream_t stream) {
    HGNN_REQUIRE(bs >= 0 && (bs == 0 || blobs) && out_dev && stage_dev && meta_host && meta_dev, "bad argument");
    std::vector<Blob> B;
    HGNN_REQUIRE(open_blobs(bs, blobs, dual, &B), "malformed graph blob (magic / field count)");
    long long layout[2 * N_KEYS];
    const DevicePlan p = plan_device(B, dual, skip_bt, layout);
    cudaStream_t s = to_stream(stream);
    // ---- graph blobs -> device staging: one DMA per run of host-adjacent blobs, or (many small scattered blobs) a host
    //      copy into the meta buffer, which goes up in one piece below
    const std::vector<long long>& blob_off = p.blob_off;
    char* stage = p.stage_in_meta ? static_cast<char*>(meta_dev) + p.blob_meta_off : static_cast<char*>(stage_dev);
    if (p.stage_in_meta) {
        char* dst = static_cast<char*>(meta_host) + p.blob_meta_off;
        for (int g = 0; g < bs; ++g) memcpy(dst + blob_off[g], B[g].base, (size_t)blob_bytes(B[g], skip_bt));
    } else {
        for (int g = 0; g < bs;) {
            int e_ = g + 1;
            while (e_ < bs && !p.run_start[e_]) ++e_;
            const size_t bytes = (size_t)(blob_off[e_ - 1] + blob_bytes(B[e_ - 1], skip_bt) - blob_off[g]);
            cudaError_t e = cudaMemcpyAsync(stage + blob_off[g], B[g].base, bytes, cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) {
                hgnn_set_error("hgnn_pack_device_upload: cudaMemcpyAsync(blobs %d..%d): %s", g, e_ - 1, cudaGetErrorString(e));
                return HGNN_ERR_CUDA;
            }
            g = e_;
        }
    }
    // ---- task table + small arrays in the meta buffer
    PackTask* tasks = static_cast<PackTask*>(meta_host);
    int* small = reinterpret_cast<int*>(static_cast<char*>(meta_host) + p.small_off);
    const int* small_dev = reinterpret_cast<const int*>(static_cast<char*>(meta_dev) + p.small_off);
    char* out = static_cast<char*>(out_dev);
    std::vector<long long> node_off(bs + 1, 0), edge_off(bs + 1, 0);
    long long nmax = 0;
    for (int g = 0; g < bs; ++g) {
        node_off[g + 1] = node_off[g] + B[g].N();
        edge_off[g + 1] = edge_off[g] + B[g].M();
        nmax = B[g].N() > nmax ? B[g].N() : nmax;
    }
    int nt = 0, ns = 0;
    auto small_task = [&](int key, int n) {   // the next n small ints -> output array `key`
        tasks[nt++] = PackTask{small_dev + ns, reinterpret_cast<int*>(out + layout[2 * key]), n, 0};
    };
    for (int k = 0; k < N_KEYS; ++k) {
        const KeyDesc& K = KEYS[k];
        if (layout[2 * k + 1] < 0) continue;
        switch (K.type) {
            case T_NODE_OFF: case T_EDGE_OFF: {
                small_task(k, bs + 1);
                const std::vector<long long>& o = K.type == T_NODE_OFF ? node_off : edge_off;
                for (int g = 0; g <= bs; ++g) small[ns++] = (int)o[g];
                break;
            }
            case T_PAD_N: {
                small_task(k, bs);
                for (int g = 0; g < bs; ++g) {
                    const float f = (float)(nmax - B[g].N());
                    memcpy(&small[ns++], &f, 4);
                }
                break;
            }
            case T_RAW: case T_RP: {
                const int drop = K.type == T_RP ? 1 : 0;
                const int kind = add_kind(k);
                int* dst = reinterpret_cast<int*>(out + layout[2 * k]);
                long long cum = 0, pos = 0;
                for (int g = 0; g < bs; ++g) {
                    const long long n = B[g].len(K.field) - drop;
                    const long long add = kind == ADD_NONE ? 0 : kind == ADD_NODE ? node_off[g] : kind == ADD_EDGE ? edge_off[g] : cum;
                    const int* src = reinterpret_cast<const int*>(stage + blob_off[g] + (B[g].ptr(K.field) - B[g].base));
                    tasks[nt++] = PackTask{src, dst + pos, (int)n, (int)add};
                    pos += n;
                    if (kind >= 0) cum += B[g].len(kind);
                }
                if (drop) {   // closing row pointer = total number of entries
                    long long tot = 0;
                    for (int g = 0; g < bs; ++g) tot += B[g].len(K.field + 1);
                    tasks[nt++] = PackTask{small_dev + ns, dst + pos, 1, 0};
                    small[ns++] = (int)tot;
                }
                break;
            }
            default: break;
        }
    }
    HGNN_REQUIRE(nt == p.n_tasks && ns == p.n_small, "internal: task count mismatch");
    cudaError_t e = cudaMemcpyAsync(meta_dev, meta_host, (size_t)p.meta_bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
        hgnn_set_error("hgnn_pack_device_upload: cudaMemcpyAsync(meta): %s", cudaGetErrorString(e));
        return HGNN_ERR_CUDA;
    }
    if (nt > 0) pack_gather_kernel<<<nt < 4 * HGNN_SM_COUNT ? nt : 4 * HGNN_SM_COUNT, 256, 0, s>>>(
        static_cast<const PackTask*>(meta_dev), nt);
    return hgnn_check_launch("hgnn_pack_device_upload");
}

// ---- padded host feature tensors of prepare_batch -------------------------------------------------------------------
// X (bs, n_feat, Nmax): X[g, f, j] = x_g[j, f] for j < N_g, zero beyond (functions/batching.py:113-127 of the reference
// builds the same zero-padded tensor with a Python loop); XL (bs, 1, Emax): the line-graph degree `dl` of every graph
// (:171), read straight from the graph blobs.  One foreign call instead of ~130 numpy slice operations per batch
// (0.12 ms of the 0.55 ms prepare_batch took on C2).  Host code only.  XL == NULL: primal batch.
extern "C" int hgnn_host_fill_features(int bs, const void* const* blobs, const float* const* x_rows, int n_feat,
                                       long long Nmax, float* X, long long Emax, float* XL) {
    HGNN_REQUIRE(bs >= 0 && blobs && x_rows && n_feat > 0 && X && Nmax >= 0, "bad argument");
    std::vector<Blob> B;
    HGNN_REQUIRE(open_blobs(bs, blobs, XL != nullptr, &B), "malformed graph blob (magic / field count)");
    for (int g = 0; g < bs; ++g) {
        const long long n = B[g].N();
        HGNN_REQUIRE(n <= Nmax && x_rows[g], "graph larger than the padded width");
        const float* src = x_rows[g];
        for (int f = 0; f < n_feat; ++f) {
            float* dst = X + ((size_t)g * n_feat + f) * Nmax;
            for (long long j = 0; j < n; ++j) dst[j] = src[j * n_feat + f];
            for (long long j = n; j < Nmax; ++j) dst[j] = 0.f;
        }
        if (XL) {
            const long long m = B[g].M();
            HGNN_REQUIRE(m <= Emax && B[g].len(F_DL) == m, "line graph larger than the padded width");
            float* dst = XL + (size_t)g * Emax;
            memcpy(dst, B[g].ptr(F_DL), (size_t)m * sizeof(float));
            for (long long j = m; j < Emax; ++j) dst[j] = 0.f;
        }
    }
    return HGNN_OK;
}
