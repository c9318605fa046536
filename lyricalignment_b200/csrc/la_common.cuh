// la_common.cuh -- shared device helpers (mbarrier / bulk-copy PTX) and kernel parameter blocks.
// sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace la {

// ---------------------------------------------------------------------------------------
// mbarrier + cp.async.bulk (TMA 1-D bulk copy, SASS UBLKCP) wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; src/dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same, with an L2 evict-first policy: for data that is streamed exactly once (the logits)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// L2 prefetch of a global range (16-byte aligned, multiple of 16 bytes)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// sub-block barrier over `nthreads` threads (whole warps). bar.sync is the .aligned form: every
// lane of a participating warp must execute it together, so reconverge first.
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------
// batch metadata shared by the kernels (device pointers, owned by the plan)
// ---------------------------------------------------------------------------------------
struct BatchMeta {
    int n_utt;
    int V;
    int mode;
    const int32_t* t_off;    // [n_utt+1] first logits row of each utterance
    const int32_t* l_off;    // [n_utt+1]
    const int32_t* labels;   // [sum L] original logit columns
    const int64_t* e_off;    // [n_utt] float offset of the emission rows in the workspace
    const int32_t* e_row;    // [n_utt] floats per emission row = round_up(L+1, 4)
    const int64_t* bp_off;   // [n_utt] uint32 offset of the packed backpointers in the workspace
    const int32_t* bp_pairs; // [n_utt] padded pair count (words per 8-frame block)
};

struct EmitParams {
    BatchMeta m;
    const float* logits;   // points at batch row `row0`
    int64_t ld;            // row stride in floats
    const float* sil;      // LOGP mode only
    int64_t ld_sil;
    float* E;              // workspace emission area
    int64_t row0;          // batch row index of logits row 0 (host path streams row ranges)
    int n_rows;
    int l2_hint;           // 1: L2 evict-first policy on the logits bulk copies (A/B knob LA_EMIT_L2_HINT)
};

struct VitParams {
    BatchMeta m;
    const float* E;
    uint32_t* bp;
    const int32_t* order;   // utterance ids of this bucket
    int n_order;
    int row_floats_max;     // widest emission row in the bucket (smem stage sizing)
    int chunk;              // frames per TMA chunk
    int stages;             // emission stages in shared memory (set by launch_viterbi)
    int ring;               // hand-off ring slots per warp boundary (set by launch_viterbi)
    int trace;              // LA_VIT_TRACE: CTA 0 records its phase clocks (perf triage)
    int32_t* first;
    int32_t* last_plus1;
    double* score;
    int32_t* status;
    double* dp_dump;        // debug/parity only: full fp64 table [T][2L+1] of a 1-utterance plan, or null
};

// N1 (la_head.cu): fused head Linear + log-sum-exp + label gather
struct HeadParams {
    const unsigned char* xp;   // packed activations [row tiles][k-steps][A_hi | A_lo]
    const unsigned char* wp;   // packed weights     [col tiles][k-steps][B_hi | B_lo]
    const float* bias;         // [V]
    float2* lse;               // [rows]: (max, sum exp(z - max)) over the softmax columns [col_lo, col_hi]
    int rows, m_tiles, n_tiles, ksteps, V, col_lo, col_hi;
    int n_splits, n_per_split; // column sweep split over n_splits CTAs per row tile (small batches); lse is [n_splits][rows]
};
struct HeadGatherParams {
    BatchMeta m;
    const float* X;            // [rows][D] activations (the Mish output)
    int64_t ldx;
    const float* W;            // [V][D] the Linear's weight
    int64_t ldw;
    const float* bias;         // [V]
    const float2* lse;         // from head_lse_kernel: [n_splits][rows]
    int n_splits;
    float* E;                  // the plan's emission area (same layout K2 writes)
    int64_t rows;
    int D;
};
size_t head_packed_bytes(int64_t rows, int D, bool weights);
int head_tile_rows();
int head_tile_cols();
cudaError_t launch_head_pack(const float* src, int64_t ld, int64_t rows, int D, void* dst, bool weights, cudaStream_t stream);
cudaError_t launch_head_lse(const HeadParams& p, int sm_count, cudaStream_t stream);
cudaError_t launch_head_gather(const HeadGatherParams& g, cudaStream_t stream);

cudaError_t launch_emit(const EmitParams& p, int sm_count, cudaStream_t stream);
cudaError_t launch_viterbi(const VitParams& p, int K, int warps, cudaStream_t stream);
cudaError_t launch_viterbi_wave(const VitParams& p, int warps, cudaStream_t stream);
int viterbi_chunk_frames(int row_floats_max);
int set_error(int code, const char* msg);
// set_error records la_last_error() and returns code

constexpr double kFloor = -10000000.0;   // utils/alignment.py:144
constexpr float kClip = -1000.0f;        // utils/alignment.py:132,134 / :18,20

}  // namespace la
