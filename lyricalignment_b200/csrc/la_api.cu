// la_api.cu -- the C ABI (include/lyricalign.h): plans, workspace layout, kernel dispatch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/lyricalign.h"
#include "la_common.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return LA_ERR_CUDA;
}
#define LA_CUDA(x)                                            \
    do {                                                      \
        cudaError_t e__ = (x);                                \
        if (e__ != cudaSuccess) return cuda_fail(e__, #x);    \
    } while (0)

namespace la {
int set_error(int code, const char* msg) { return fail(code, msg); }
void logmel_release_tables();
}

// K3 launch shapes. Buckets 0-2: the wavefront kernel (la_viterbi_wave.cuh): one warp per utterance up to 63 pairs
// (one pair per lane up to 32 pairs, two from 33 on -- chosen per utterance inside ONE launch), then 64 columns per
// warp on up to 4 / up to 10 warps. The other buckets: the row-synchronous pipeline kernel, {pairs per lane, max
// warps}. One CTA per utterance; a launch is sized for the widest utterance of its bucket, surplus warps exit.
constexpr int kBuckets = 7;
constexpr int kWaveBuckets = 3;
struct BucketShape { int K; int max_warps; };
static const BucketShape kShape[kBuckets] = {{0, 1}, {2, 4}, {2, 10}, {2, 16}, {2, 32}, {4, 32}, {8, 32}};
static int bucket_for_pairs(int pairs) {
    if (pairs <= 63) return 0;
    for (int b = 1; b < kBuckets; ++b)                        // wave shapes: pair i lives in column i + 1
        if (pairs + (b < kWaveBuckets ? 1 : 0) <= 32 * kShape[b].K * kShape[b].max_warps) return b;
    return kBuckets - 1;
}

struct la_plan {
    int mode = 0, n_utt = 0, V = 0, device = 0, sm_count = 148;
    int64_t total_T = 0, total_L = 0;
    std::vector<int32_t> t_off, l_off, e_row, bp_pairs;
    std::vector<int64_t> e_off, bp_off;     // floats / uint32 words, relative to their area
    size_t emit_bytes = 0, bp_bytes = 0;
    std::vector<int32_t> order[kBuckets];
    int row_max[kBuckets] = {0};
    int warps_max[kBuckets] = {0};
    // device metadata blob
    void* d_meta = nullptr;
    bool meta_pooled = false;
    la::BatchMeta meta{};
    const int32_t* d_order[kBuckets] = {nullptr};
    // the last stream work of this plan was enqueued on: la_plan_destroy() records an event there so the
    // pooled metadata block is not handed to another plan while K2/K3 may still be reading it
    mutable cudaStream_t last_stream = nullptr;
    mutable bool used = false;
};

// Host-path context (la_align_host): one per device, grown on demand, reused across plans so a
// per-clip call (the reference's batch size 1) pays no cudaMalloc after the first.
struct HostCtx {
    std::mutex mu;
    void* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    void* d_out = nullptr;
    size_t out_bytes = 0;
    void* h_out = nullptr;      // pinned bounce buffer for the results
    size_t h_out_bytes = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
};
static HostCtx g_host[64];

// Plan metadata lives in pooled 64 KiB device blocks: cudaMalloc / cudaFree (the latter a device-wide
// sync) cost milliseconds next to a 40 MB clip, and the reference's entry points decode one clip
// per call.
constexpr size_t kMetaBlock = 64 << 10;
struct PoolBlock { void* p; cudaEvent_t busy; };   // busy == nullptr: free right away
struct DevInfo {
    std::mutex mu;
    int sm_count = 0;
    std::vector<PoolBlock> free_blocks;
};
static DevInfo g_dev[64];

static int meta_alloc(int device, size_t bytes, void** out, bool* pooled) {
    DevInfo& D = g_dev[device];
    if (bytes <= kMetaBlock) {
        std::lock_guard<std::mutex> lock(D.mu);
        *pooled = true;
        for (size_t i = D.free_blocks.size(); i-- > 0;) {
            PoolBlock& b = D.free_blocks[i];
            if (b.busy) {
                if (cudaEventQuery(b.busy) != cudaSuccess) { cudaGetLastError(); continue; }   // its last user is still running
                cudaEventDestroy(b.busy);
            }
            *out = b.p;
            D.free_blocks.erase(D.free_blocks.begin() + (long)i);
            return LA_OK;
        }
        LA_CUDA(cudaMalloc(out, kMetaBlock));
        return LA_OK;
    }
    *pooled = false;
    LA_CUDA(cudaMalloc(out, bytes));
    return LA_OK;
}
// `stream`/`used`: where the plan's kernels were last enqueued. A pooled block goes back with an event
// recorded there and is only reused once that event has completed; cudaFree synchronises by itself.
static void meta_free(int device, void* p, bool pooled, cudaStream_t stream = nullptr, bool used = false) {
    if (!p) return;
    if (pooled) {
        cudaEvent_t ev = nullptr;
        if (used) {
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventRecord(ev, stream) != cudaSuccess) {
                cudaGetLastError();
                if (ev) cudaEventDestroy(ev);
                ev = nullptr;
                cudaDeviceSynchronize();          // could not fence the block: wait for everything instead
            }
        }
        std::lock_guard<std::mutex> lock(g_dev[device].mu);
        g_dev[device].free_blocks.push_back(PoolBlock{p, ev});
    } else {
        cudaFree(p);
    }
}
static int device_sm_count(int device, int* out) {
    DevInfo& D = g_dev[device];
    std::lock_guard<std::mutex> lock(D.mu);
    if (D.sm_count == 0) {
        int n = 0;
        LA_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
        D.sm_count = n;
    }
    *out = D.sm_count;
    return LA_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" {

const char* la_version(void) { return "lyricalign-b200 0.1 (sm_100a)"; }
const char* la_last_error(void) { return g_err.c_str(); }

// Releases everything the library caches between calls (see lyricalign.h "Cached state"): the host-path
// staging context of every device, the pooled plan-metadata blocks, K1's basis tables. Plans that are
// still alive stay valid (their metadata blocks are theirs until la_plan_destroy).
void la_shutdown(void) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    for (int d = 0; d < 64; ++d) {
        HostCtx& C = g_host[d];
        {
            std::lock_guard<std::mutex> lock(C.mu);
            if (C.s_copy || C.d_ws || C.d_out || C.h_out || C.d_stage[0]) {
                cudaSetDevice(d);
                if (C.s_copy) cudaStreamSynchronize(C.s_copy);
                if (C.s_comp) cudaStreamSynchronize(C.s_comp);
                for (int i = 0; i < 2; ++i) {
                    if (C.d_stage[i]) cudaFree(C.d_stage[i]);
                    if (C.ev_copied[i]) cudaEventDestroy(C.ev_copied[i]);
                    if (C.ev_done[i]) cudaEventDestroy(C.ev_done[i]);
                    C.d_stage[i] = nullptr; C.ev_copied[i] = nullptr; C.ev_done[i] = nullptr;
                }
                if (C.d_ws) cudaFree(C.d_ws);
                if (C.d_out) cudaFree(C.d_out);
                if (C.h_out) cudaFreeHost(C.h_out);
                if (C.s_copy) cudaStreamDestroy(C.s_copy);
                if (C.s_comp) cudaStreamDestroy(C.s_comp);
                C.d_ws = C.d_out = C.h_out = nullptr;
                C.s_copy = C.s_comp = nullptr;
                C.stage_bytes = C.ws_bytes = C.out_bytes = C.h_out_bytes = 0;
            }
        }
        DevInfo& D = g_dev[d];
        std::lock_guard<std::mutex> lock(D.mu);
        if (!D.free_blocks.empty()) {
            cudaSetDevice(d);
            for (PoolBlock& b : D.free_blocks) {
                if (b.busy) { cudaEventSynchronize(b.busy); cudaEventDestroy(b.busy); }
                cudaFree(b.p);
            }
            D.free_blocks.clear();
        }
    }
    la::logmel_release_tables();
    cudaSetDevice(cur);
    cudaGetLastError();
}
int la_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int plan_create_impl(la_plan** out, int mode, int n_utt, int V, const int32_t* h_t_len,
                            const int32_t* h_l_len, const int32_t* h_labels, int device);

int la_plan_create(la_plan** out, int mode, int n_utt, int V, const int32_t* h_t_len,
                   const int32_t* h_l_len, const int32_t* h_labels, int device) {
    try {                                    // nothing may throw across the C ABI
        return plan_create_impl(out, mode, n_utt, V, h_t_len, h_l_len, h_labels, device);
    } catch (const std::bad_alloc&) {
        return fail(LA_ERR_ALLOC, "out of host memory");
    } catch (...) {
        return fail(LA_ERR_ARG, "unexpected exception in la_plan_create");
    }
}

}  // extern "C"

static int plan_create_impl(la_plan** out, int mode, int n_utt, int V, const int32_t* h_t_len,
                            const int32_t* h_l_len, const int32_t* h_labels, int device) {
    if (!out || n_utt < 0 || (n_utt > 0 && (!h_t_len || !h_l_len))) return fail(LA_ERR_ARG, "null argument");
    if (mode < 0 || mode > 2) return fail(LA_ERR_ARG, "mode must be LA_MODE_CTC/CE/LOGP");
    if ((mode == LA_MODE_CTC && V < 3) || V < 1) return fail(LA_ERR_ARG, "V too small for this mode");
    la_plan* P = new (std::nothrow) la_plan();
    if (!P) return fail(LA_ERR_ALLOC, "out of host memory");
    P->mode = mode; P->n_utt = n_utt; P->V = V; P->device = device;
    P->t_off.assign(n_utt + 1, 0);
    P->l_off.assign(n_utt + 1, 0);
    P->e_row.assign(n_utt, 0);
    P->bp_pairs.assign(n_utt, 0);
    P->e_off.assign(n_utt, 0);
    P->bp_off.assign(n_utt, 0);
    int64_t tsum = 0, lsum = 0;
    for (int u = 0; u < n_utt; ++u) {
        if (h_t_len[u] < 0 || h_l_len[u] < 0) { delete P; return fail(LA_ERR_ARG, "negative length"); }
        if (h_l_len[u] > LA_MAX_LABELS) { delete P; return fail(LA_ERR_LIMIT, "label row longer than LA_MAX_LABELS"); }
        tsum += h_t_len[u]; lsum += h_l_len[u];
        if (tsum > INT32_MAX || lsum > INT32_MAX) { delete P; return fail(LA_ERR_LIMIT, "batch too large for int32 offsets"); }
        P->t_off[u + 1] = (int32_t)tsum;
        P->l_off[u + 1] = (int32_t)lsum;
    }
    P->total_T = tsum; P->total_L = lsum;
    if (lsum > 0 && !h_labels) { delete P; return fail(LA_ERR_ARG, "null labels"); }
    for (int64_t i = 0; i < lsum; ++i)
        if (h_labels[i] < 0 || h_labels[i] >= V) { delete P; return fail(LA_ERR_ARG, "label column out of range"); }

    // buckets + workspace layout
    std::vector<int> bucket_of(n_utt, 0);
    for (int u = 0; u < n_utt; ++u) {
        const int L = h_l_len[u];
        const int b = bucket_for_pairs(L + 1);
        bucket_of[u] = b;
        P->e_row[u] = (int32_t)align_up((size_t)L + 1, 4);
        P->row_max[b] = std::max(P->row_max[b], P->e_row[u]);
        P->order[b].push_back(u);
    }
    // longest utterance first: one CTA per utterance, dispatched in grid order, so the tail of a launch is its
    // shortest clips rather than whichever long one happened to come last
    for (int b = 0; b < kBuckets; ++b)
        std::stable_sort(P->order[b].begin(), P->order[b].end(),
                         [&](int32_t x, int32_t y) { return h_t_len[x] > h_t_len[y]; });
    size_t e_floats = 0, bp_words = 0;
    for (int u = 0; u < n_utt; ++u) {
        const BucketShape& sh = kShape[bucket_of[u]];
        P->e_off[u] = (int64_t)e_floats;
        e_floats += (size_t)h_t_len[u] * P->e_row[u];
        P->bp_off[u] = (int64_t)bp_words;
        if (bucket_of[u] < kWaveBuckets) {                         // wavefront kernel: pair i lives in column i + 1 from 33 pairs on
            const int pairs = h_l_len[u] + 1;
            const int warps = bucket_of[u] == 0 ? 1 : (pairs + 1 + 63) / 64;
            P->warps_max[bucket_of[u]] = std::max(P->warps_max[bucket_of[u]], warps);
            P->bp_pairs[u] = bucket_of[u] == 0 ? (pairs <= 32 ? 32 : 64) : 64 * warps;
            bp_words += (size_t)((h_t_len[u] + 7) / 8) * P->bp_pairs[u];
        } else {
            const int warps = std::max(1, (h_l_len[u] + 1 + 32 * sh.K - 1) / (32 * sh.K));
            P->warps_max[bucket_of[u]] = std::max(P->warps_max[bucket_of[u]], warps);
            P->bp_pairs[u] = 32 * warps * sh.K;
            bp_words += (size_t)((h_t_len[u] + 7) / 8) * P->bp_pairs[u];
        }
    }
    P->emit_bytes = align_up(e_floats * 4 + 16, 256);
    P->bp_bytes = align_up(bp_words * 4 + 16, 256);

    // device metadata blob
    if (device < 0 || device >= 64) { delete P; return fail(LA_ERR_ARG, "device index out of range"); }
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) { delete P; return cuda_fail(ce, "cudaSetDevice"); }
    if (int rc = device_sm_count(device, &P->sm_count)) { delete P; return rc; }
    std::vector<unsigned char> blob;
    auto put = [&](const void* src, size_t bytes) -> size_t {
        const size_t off = align_up(blob.size(), 16);
        blob.resize(off + std::max<size_t>(bytes, 16));
        if (bytes) memcpy(blob.data() + off, src, bytes);
        return off;
    };
    const size_t o_t = put(P->t_off.data(), P->t_off.size() * 4);
    const size_t o_l = put(P->l_off.data(), P->l_off.size() * 4);
    const size_t o_lab = put(h_labels, (size_t)lsum * 4);
    const size_t o_eo = put(P->e_off.data(), P->e_off.size() * 8);
    const size_t o_er = put(P->e_row.data(), P->e_row.size() * 4);
    const size_t o_bo = put(P->bp_off.data(), P->bp_off.size() * 8);
    const size_t o_bpp = put(P->bp_pairs.data(), P->bp_pairs.size() * 4);
    size_t o_ord[kBuckets];
    for (int b = 0; b < kBuckets; ++b) o_ord[b] = put(P->order[b].data(), P->order[b].size() * 4);
    if (int rc = meta_alloc(device, blob.size(), &P->d_meta, &P->meta_pooled)) { delete P; return rc; }
    ce = cudaMemcpy(P->d_meta, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { meta_free(device, P->d_meta, P->meta_pooled); delete P; return cuda_fail(ce, "cudaMemcpy(meta)"); }
    unsigned char* d = static_cast<unsigned char*>(P->d_meta);
    P->meta.n_utt = n_utt; P->meta.V = V; P->meta.mode = mode;
    P->meta.t_off = reinterpret_cast<const int32_t*>(d + o_t);
    P->meta.l_off = reinterpret_cast<const int32_t*>(d + o_l);
    P->meta.labels = reinterpret_cast<const int32_t*>(d + o_lab);
    P->meta.e_off = reinterpret_cast<const int64_t*>(d + o_eo);
    P->meta.e_row = reinterpret_cast<const int32_t*>(d + o_er);
    P->meta.bp_off = reinterpret_cast<const int64_t*>(d + o_bo);
    P->meta.bp_pairs = reinterpret_cast<const int32_t*>(d + o_bpp);
    for (int b = 0; b < kBuckets; ++b) P->d_order[b] = reinterpret_cast<const int32_t*>(d + o_ord[b]);
    *out = P;
    return LA_OK;
}

extern "C" {

void la_plan_destroy(la_plan* P) {
    if (!P) return;
    int cur = -1;
    if (P->used && cudaGetDevice(&cur) == cudaSuccess && cur != P->device) cudaSetDevice(P->device); else cur = -1;
    meta_free(P->device, P->d_meta, P->meta_pooled, P->last_stream, P->used);
    if (cur >= 0) cudaSetDevice(cur);
    delete P;
}

size_t la_plan_workspace_bytes(const la_plan* P) { return P ? P->emit_bytes + P->bp_bytes : 0; }
int64_t la_plan_total_frames(const la_plan* P) { return P ? P->total_T : 0; }
int64_t la_plan_total_labels(const la_plan* P) { return P ? P->total_L : 0; }
int la_plan_num_launches(const la_plan* P) {
    if (!P) return 0;
    int n = P->total_T > 0 ? 1 : 0;
    for (int b = 0; b < kBuckets; ++b) n += P->order[b].empty() ? 0 : 1;
    return n;
}

int la_plan_utt_layout(const la_plan* P, int utt, int64_t* emit_off_bytes, int32_t* row_floats,
                       int64_t* bp_off_bytes, int32_t* pairs_padded) {
    if (!P || utt < 0 || utt >= P->n_utt) return fail(LA_ERR_ARG, "bad utterance index");
    if (emit_off_bytes) *emit_off_bytes = P->e_off[utt] * 4;
    if (row_floats) *row_floats = P->e_row[utt];
    if (bp_off_bytes) *bp_off_bytes = (int64_t)P->emit_bytes + P->bp_off[utt] * 4;
    if (pairs_padded) *pairs_padded = P->bp_pairs[utt];
    return LA_OK;
}

int la_plan_utt_bp_layout(const la_plan* P, int utt, int32_t* word_rows, int32_t* col_shift) {
    if (!P || utt < 0 || utt >= P->n_utt) return fail(LA_ERR_ARG, "bad utterance index");
    const int T = P->t_off[utt + 1] - P->t_off[utt];
    const int L = P->l_off[utt + 1] - P->l_off[utt];
    if (word_rows) *word_rows = (T + 7) / 8;
    if (col_shift) *col_shift = (bucket_for_pairs(L + 1) < kWaveBuckets && L + 1 > 32) ? 1 : 0;
    return LA_OK;
}

// rows [row0, row0 + n_rows) of the batch; d_logits points at row `row0`
static int emit_rows(const la_plan* P, const float* d_logits, int64_t ld, const float* d_sil,
                     int64_t ld_sil, void* d_ws, int64_t row0, int64_t n_rows, cudaStream_t stream);

int la_emit(const la_plan* P, const float* d_logits, int64_t ld, const float* d_sil, int64_t ld_sil,
            void* d_ws, void* stream) {
    if (!P || !d_ws) return fail(LA_ERR_ARG, "null argument");
    if (P->total_T == 0) return LA_OK;
    if (!d_logits) return fail(LA_ERR_ARG, "null logits");
    if (ld < P->V) return fail(LA_ERR_ARG, "row stride smaller than V");
    if (P->mode == LA_MODE_LOGP && !d_sil) return fail(LA_ERR_ARG, "LA_MODE_LOGP needs the silence column");
    if (reinterpret_cast<uintptr_t>(d_logits) & 15) return fail(LA_ERR_ARG, "logits base must be 16-byte aligned");
    LA_CUDA(cudaSetDevice(P->device));
    return emit_rows(P, d_logits, ld, d_sil, ld_sil, d_ws, 0, P->total_T, static_cast<cudaStream_t>(stream));
}

static int viterbi_impl(const la_plan* P, void* d_ws, int32_t* d_first, int32_t* d_last, double* d_score,
                        int32_t* d_status, double* d_dp, void* stream) {
    if (!P || !d_ws || !d_score || !d_status) return fail(LA_ERR_ARG, "null argument");
    if (P->total_L > 0 && (!d_first || !d_last)) return fail(LA_ERR_ARG, "null output");
    LA_CUDA(cudaSetDevice(P->device));
    for (int b = 0; b < kBuckets; ++b) {
        if (P->order[b].empty()) continue;
        la::VitParams vp;
        vp.m = P->meta;
        vp.E = static_cast<const float*>(d_ws);
        vp.bp = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(d_ws) + P->emit_bytes);
        vp.order = P->d_order[b];
        vp.n_order = (int)P->order[b].size();
        vp.row_floats_max = P->row_max[b];
        vp.chunk = la::viterbi_chunk_frames(P->row_max[b]);
        vp.first = d_first; vp.last_plus1 = d_last; vp.score = d_score; vp.status = d_status;
        vp.dp_dump = d_dp;
        if (b < kWaveBuckets) LA_CUDA(la::launch_viterbi_wave(vp, P->warps_max[b], static_cast<cudaStream_t>(stream)));
        else LA_CUDA(la::launch_viterbi(vp, kShape[b].K, P->warps_max[b], static_cast<cudaStream_t>(stream)));
    }
    P->last_stream = static_cast<cudaStream_t>(stream);
    P->used = true;
    return LA_OK;
}

int la_viterbi(const la_plan* P, void* d_ws, int32_t* d_first, int32_t* d_last, double* d_score,
               int32_t* d_status, void* stream) {
    return viterbi_impl(P, d_ws, d_first, d_last, d_score, d_status, nullptr, stream);
}

int la_viterbi_debug(const la_plan* P, void* d_ws, int32_t* d_first, int32_t* d_last, double* d_score,
                     int32_t* d_status, double* d_dp, void* stream) {
    if (P && P->n_utt != 1) return fail(LA_ERR_ARG, "la_viterbi_debug takes a 1-utterance plan");
    return viterbi_impl(P, d_ws, d_first, d_last, d_score, d_status, d_dp, stream);
}

int la_align(const la_plan* P, const float* d_logits, int64_t ld, void* d_ws, int32_t* d_first,
             int32_t* d_last, double* d_score, int32_t* d_status, void* stream) {
    if (P && P->mode == LA_MODE_LOGP) return fail(LA_ERR_ARG, "la_align needs LA_MODE_CTC or LA_MODE_CE");
    int rc = la_emit(P, d_logits, ld, nullptr, 0, d_ws, stream);
    if (rc) return rc;
    return la_viterbi(P, d_ws, d_first, d_last, d_score, d_status, stream);
}

// ---- N1: fused head (Linear + log-softmax + gather) ------------------------------------------
static inline bool head_dim_ok(int D) { return D >= 32 && D <= 1024 && D % 32 == 0; }

size_t la_head_packed_weight_bytes(int V, int D) {
    if (V <= 0 || !head_dim_ok(D)) return 0;
    return la::head_packed_bytes(V, D, true);
}

int la_head_pack_weights(const float* d_W, int64_t ldw, int V, int D, void* d_packed, void* stream) {
    if (!d_W || !d_packed) return fail(LA_ERR_ARG, "null argument");
    if (V <= 0 || !head_dim_ok(D)) return fail(LA_ERR_ARG, "D must be a multiple of 32 in [32, 1024]");
    if (ldw < D || (ldw & 3) || (reinterpret_cast<uintptr_t>(d_W) & 15)) return fail(LA_ERR_ARG, "weight rows must be 16-byte aligned");
    LA_CUDA(la::launch_head_pack(d_W, ldw, V, D, d_packed, true, static_cast<cudaStream_t>(stream)));
    return LA_OK;
}

// fewer row tiles than SMs (per-clip calls): split the column sweep so the whole chip works on it
static int head_splits(const la_plan* P, int* per_split) {
    const int m_tiles = (int)((P->total_T + la::head_tile_rows() - 1) / la::head_tile_rows());
    const int n_tiles = (P->V + la::head_tile_cols() - 1) / la::head_tile_cols();
    int want = 1;
    if (m_tiles > 0 && m_tiles < P->sm_count) want = std::min(n_tiles, (2 * P->sm_count + m_tiles - 1) / m_tiles);
    const int per = (n_tiles + want - 1) / want;
    if (per_split) *per_split = per;
    return (n_tiles + per - 1) / per;
}

size_t la_head_workspace_bytes(const la_plan* P, int D) {
    if (!P || !head_dim_ok(D)) return 0;
    return align_up(la::head_packed_bytes(P->total_T, D, false), 256) +
           align_up((size_t)std::max<int64_t>(P->total_T, 1) * 8 * head_splits(P, nullptr), 256);
}

int la_head_emit(const la_plan* P, const float* d_X, int64_t ldx, int D, const float* d_W, int64_t ldw,
                 const float* d_bias, const void* d_packed_w, void* d_head_ws, void* d_ws, void* stream) {
    if (!P || !d_ws) return fail(LA_ERR_ARG, "null argument");
    if (P->mode != LA_MODE_CTC && P->mode != LA_MODE_CE) return fail(LA_ERR_ARG, "la_head_emit needs LA_MODE_CTC or LA_MODE_CE");
    if (P->total_T == 0) return LA_OK;
    if (!d_X || !d_W || !d_bias || !d_packed_w || !d_head_ws) return fail(LA_ERR_ARG, "null argument");
    if (!head_dim_ok(D)) return fail(LA_ERR_ARG, "D must be a multiple of 32 in [32, 1024]");
    if (ldx < D || (ldx & 3) || (reinterpret_cast<uintptr_t>(d_X) & 15)) return fail(LA_ERR_ARG, "activation rows must be 16-byte aligned");
    if (ldw < D) return fail(LA_ERR_ARG, "weight row stride smaller than D");
    LA_CUDA(cudaSetDevice(P->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* hw = static_cast<unsigned char*>(d_head_ws);
    const size_t xp_bytes = align_up(la::head_packed_bytes(P->total_T, D, false), 256);
    LA_CUDA(la::launch_head_pack(d_X, ldx, P->total_T, D, hw, false, st));
    la::HeadParams hp;
    hp.xp = hw;
    hp.wp = static_cast<const unsigned char*>(d_packed_w);
    hp.bias = d_bias;
    hp.lse = reinterpret_cast<float2*>(hw + xp_bytes);
    hp.rows = (int)P->total_T;
    hp.m_tiles = (int)((P->total_T + la::head_tile_rows() - 1) / la::head_tile_rows());
    hp.n_tiles = (P->V + la::head_tile_cols() - 1) / la::head_tile_cols();
    hp.ksteps = D / 16;
    hp.V = P->V;
    hp.col_lo = P->mode == LA_MODE_CTC ? 1 : 0;            // utils/alignment.py:123: softmax over [:, :, 1:-1]
    hp.col_hi = P->mode == LA_MODE_CTC ? P->V - 2 : P->V - 1;
    hp.n_splits = head_splits(P, &hp.n_per_split);
    LA_CUDA(la::launch_head_lse(hp, P->sm_count, st));
    la::HeadGatherParams gp;
    gp.m = P->meta;
    gp.X = d_X; gp.ldx = ldx; gp.W = d_W; gp.ldw = ldw; gp.bias = d_bias;
    gp.lse = hp.lse;
    gp.n_splits = hp.n_splits;
    gp.E = static_cast<float*>(d_ws);
    gp.rows = P->total_T;
    gp.D = D;
    LA_CUDA(la::launch_head_gather(gp, st));
    P->last_stream = st;
    P->used = true;
    return LA_OK;
}

int la_head_align(const la_plan* P, const float* d_X, int64_t ldx, int D, const float* d_W, int64_t ldw,
                  const float* d_bias, const void* d_packed_w, void* d_head_ws, void* d_ws, int32_t* d_first,
                  int32_t* d_last, double* d_score, int32_t* d_status, void* stream) {
    int rc = la_head_emit(P, d_X, ldx, D, d_W, ldw, d_bias, d_packed_w, d_head_ws, d_ws, stream);
    if (rc) return rc;
    return la_viterbi(P, d_ws, d_first, d_last, d_score, d_status, stream);
}

static int grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return LA_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    LA_CUDA(cudaMalloc(p, need));
    *have = need;
    return LA_OK;
}

int la_align_host(la_plan* P, const float* h_logits, int64_t ld, int32_t* h_first, int32_t* h_last,
                  double* h_score, int32_t* h_status, size_t staging_bytes) {
    if (!P || !h_score || !h_status) return fail(LA_ERR_ARG, "null argument");
    if (P->mode == LA_MODE_LOGP) return fail(LA_ERR_ARG, "la_align_host needs LA_MODE_CTC or LA_MODE_CE");
    if (P->total_T > 0 && !h_logits) return fail(LA_ERR_ARG, "null logits");
    if (P->total_L > 0 && (!h_first || !h_last)) return fail(LA_ERR_ARG, "null output");
    if (ld < P->V) return fail(LA_ERR_ARG, "row stride smaller than V");
    if (P->device < 0 || P->device >= 64) return fail(LA_ERR_ARG, "device index out of range");
    LA_CUDA(cudaSetDevice(P->device));
    HostCtx& C = g_host[P->device];
    std::lock_guard<std::mutex> lock(C.mu);
    const size_t row_bytes = (size_t)ld * 4;
    // default staging: 64 MiB. A single clip (21-64 MB of logits) then goes over in ONE copy followed by ONE K2 launch
    // -- K2 runs at 6.5 TB/s, so overlapping it with a 50 GB/s copy buys nothing and every extra chunk costs an
    // event round trip (round 1 used 16 MiB chunks: 42 GB/s effective per clip); big batches stream in 64 MiB chunks.
    if (staging_bytes == 0) staging_bytes = (size_t)64 << 20;
    size_t rows_per_stage = std::max<size_t>(1, staging_bytes / row_bytes);
    rows_per_stage = std::min<size_t>(rows_per_stage, (size_t)std::max<int64_t>(P->total_T, 1));
    if (!C.s_copy) {
        LA_CUDA(cudaStreamCreateWithFlags(&C.s_copy, cudaStreamNonBlocking));
        LA_CUDA(cudaStreamCreateWithFlags(&C.s_comp, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            LA_CUDA(cudaEventCreateWithFlags(&C.ev_copied[i], cudaEventDisableTiming));
            LA_CUDA(cudaEventCreateWithFlags(&C.ev_done[i], cudaEventDisableTiming));
        }
    }
    const size_t need = align_up(rows_per_stage * row_bytes + 16, 256);
    if (C.stage_bytes < need) {
        for (int i = 0; i < 2; ++i) {
            if (C.d_stage[i]) cudaFree(C.d_stage[i]);
            C.d_stage[i] = nullptr;
        }
        C.stage_bytes = 0;
        for (int i = 0; i < 2; ++i) LA_CUDA(cudaMalloc(&C.d_stage[i], need));
        C.stage_bytes = need;
    }
    const size_t lab_bytes = align_up((size_t)P->total_L * 4, 16);
    const size_t sc_bytes = align_up((size_t)P->n_utt * 8, 16);
    const size_t st_bytes = align_up((size_t)P->n_utt * 4, 16);
    const size_t out_bytes = 2 * lab_bytes + sc_bytes + st_bytes + 64;
    int rc = grow(&C.d_ws, &C.ws_bytes, la_plan_workspace_bytes(P));
    if (rc) return rc;
    rc = grow(&C.d_out, &C.out_bytes, out_bytes);
    if (rc) return rc;
    if (C.h_out_bytes < out_bytes) {
        if (C.h_out) cudaFreeHost(C.h_out);
        C.h_out = nullptr; C.h_out_bytes = 0;
        LA_CUDA(cudaMallocHost(&C.h_out, out_bytes));
        C.h_out_bytes = out_bytes;
    }
    unsigned char* o = static_cast<unsigned char*>(C.d_out);
    int32_t* d_first = reinterpret_cast<int32_t*>(o);
    int32_t* d_last = reinterpret_cast<int32_t*>(o + lab_bytes);
    double* d_score = reinterpret_cast<double*>(o + 2 * lab_bytes);
    int32_t* d_status = reinterpret_cast<int32_t*>(o + 2 * lab_bytes + sc_bytes);

    // chunked H2D (copy stream) overlapped with K2 (compute stream), two staging buffers
    int64_t row = 0;
    int i = 0;
    bool used[2] = {false, false};
    while (row < P->total_T) {
        const int64_t n = std::min<int64_t>((int64_t)rows_per_stage, P->total_T - row);
        const int sbuf = i & 1;
        // a failure must not leave queued work reading the staging buffers the next call will overwrite
#define LA_CUDA_DRAIN(x)                                                                            \
        do {                                                                                        \
            cudaError_t e__ = (x);                                                                  \
            if (e__ != cudaSuccess) { cudaStreamSynchronize(C.s_copy); cudaStreamSynchronize(C.s_comp); return cuda_fail(e__, #x); } \
        } while (0)
        if (used[sbuf]) LA_CUDA_DRAIN(cudaStreamWaitEvent(C.s_copy, C.ev_done[sbuf], 0));
        LA_CUDA_DRAIN(cudaMemcpyAsync(C.d_stage[sbuf], h_logits + row * ld, (size_t)n * row_bytes,
                                      cudaMemcpyHostToDevice, C.s_copy));
        LA_CUDA_DRAIN(cudaEventRecord(C.ev_copied[sbuf], C.s_copy));
        LA_CUDA_DRAIN(cudaStreamWaitEvent(C.s_comp, C.ev_copied[sbuf], 0));
        rc = emit_rows(P, static_cast<const float*>(C.d_stage[sbuf]), ld, nullptr, 0, C.d_ws, row, n, C.s_comp);
        if (rc) { cudaStreamSynchronize(C.s_copy); cudaStreamSynchronize(C.s_comp); return rc; }
        LA_CUDA_DRAIN(cudaEventRecord(C.ev_done[sbuf], C.s_comp));
        used[sbuf] = true;
        row += n;
        ++i;
    }
    rc = la_viterbi(P, C.d_ws, d_first, d_last, d_score, d_status, C.s_comp);
    if (rc) { cudaStreamSynchronize(C.s_copy); cudaStreamSynchronize(C.s_comp); return rc; }
    // one D2H of the packed results into the pinned bounce buffer, then scatter on the host
    LA_CUDA_DRAIN(cudaMemcpyAsync(C.h_out, C.d_out, out_bytes - 64, cudaMemcpyDeviceToHost, C.s_comp));
    LA_CUDA_DRAIN(cudaStreamSynchronize(C.s_comp));
#undef LA_CUDA_DRAIN
    const unsigned char* h = static_cast<const unsigned char*>(C.h_out);
    if (P->total_L > 0) {
        memcpy(h_first, h, (size_t)P->total_L * 4);
        memcpy(h_last, h + lab_bytes, (size_t)P->total_L * 4);
    }
    if (P->n_utt > 0) {
        memcpy(h_score, h + 2 * lab_bytes, (size_t)P->n_utt * 8);
        memcpy(h_status, h + 2 * lab_bytes + sc_bytes, (size_t)P->n_utt * 4);
    }
    return LA_OK;
}

}  // extern "C"

static int emit_rows(const la_plan* P, const float* d_logits, int64_t ld, const float* d_sil,
                     int64_t ld_sil, void* d_ws, int64_t row0, int64_t n_rows, cudaStream_t stream) {
    la::EmitParams ep;
    ep.m = P->meta;
    ep.logits = d_logits; ep.ld = ld; ep.sil = d_sil; ep.ld_sil = ld_sil;
    ep.E = static_cast<float*>(d_ws);
    ep.row0 = row0;
    { static const int hint = [] { const char* e = getenv("LA_EMIT_L2_HINT"); return e ? atoi(e) : 0; }(); ep.l2_hint = hint; }
    ep.n_rows = (int)n_rows;
    LA_CUDA(la::launch_emit(ep, P->sm_count, stream));
    P->last_stream = stream;
    P->used = true;
    return LA_OK;
}
