// la_logmel.cu -- K1: Whisper-style 80-bin log-mel front end as a framed, windowed DFT on the
// 5th-gen tensor cores (tcgen05.mma kind::f16, operands sliced on a fixed grid so that the leading
// product chain is EXACT in the fp32 TMEM accumulator).
//
// Replaces whisper.audio.log_mel_spectrogram as called at module/align_model.py:84 (the
// reference runs it on the CPU through torch.stft): hann(400, periodic) window, hop 160,
// centre=True reflect padding, last frame dropped, |.|^2, 80 x 201 Slaney mel filterbank,
// log10(clamp 1e-10), max(x, GLOBAL max - 8), (x + 4) / 4.
//
// Formulation. The Hann window is symmetric (w[n] = w[400-n], w[0] = 0), so with the folded
// inputs  e[n] = x[n] + x[400-n],  o[n] = x[n] - x[400-n]  (n = 0..207; rows 0 and 201.. of the
// basis are zero, row 200 of the cosine basis is halved because e[200] = 2 x[200]):
//     Re X[k] = sum_n e[n] * w[n] cos(2 pi k n / 400)        (pass 0)
//     Im X[k] = sum_n o[n] * (-w[n] sin(2 pi k n / 400))     (pass 1)
// i.e. two GEMMs [128 frames x 208] x [208 x 208] per tile, run one after the other on the same
// pair of TMEM accumulators.
//
// Precision (round 2). The tensor core's fp32 accumulate TRUNCATES (scripts/emulate_k1_precision.py),
// so round 1's 75-deep 3xTF32 chain carried a biased error proportional to the PARTIAL sums, which for
// a weak bin next to a strong one are far larger than the result (worst cell 3.3e-4 in log10 units).
// Here every operand is sliced on a fixed grid instead (scripts/emulate_k1_precision_r2.py):
//     T  = fold * 2^14 / S                S = power of two > 2 max|x| over the tile, so |T| < 2^14
//     A1 = T rounded to a multiple of 64  (8 significant bits: exact in fp16)
//     A2 = fp16(T - A1),  A3 = fp16(T - A1 - A2)          (11 bits each; |A3| < 2^-7, mostly normal)
// and the same for the basis (B1, B2, B3 from 2^14 w[n] cos|sin, sliced on the host in fp64). Then
//     Acc0 = sum A1*B1                      each product is 2^12 x an integer < 2^16 and there are 208 of
//                                           them: every partial sum has < 24 significant bits, so the chain
//                                           is EXACT in fp32 whatever the rounding mode;
//     Acc1 = sum A3*B1 + A1*B3 + A2*B2 + A2*B1 + A1*B2      terms <= 2^-9 of Acc0's: truncation there
//                                           is 2^-9 of what it was;
//     X    = 2^-28 S (Acc0 + Acc1)          one fp32 round-to-nearest add in the epilogue.
// Dropped: A2*B3 + A3*B2 + A3*B3 (2^-30 of full scale). kind::f16 has K = 16 at twice the TF32 rate, so
// the 6 x 13 MMAs per pass cost the tensor pipe what round 1's 3 x 25 did. Emulated worst cell
// 6e-6 (the reference's own fp32 torch.stft: 3e-5 .. 8e-5), measured: profiles/logmel_precision_r2.txt.
//
// One persistent CTA per SM (896 threads), 128-frame tiles, warp-specialised:
//   warp 0       producer: 1-D TMA bulk copies of the tile's waveform (17 copies of 8 hop rows into a
//                staging area padded by 8 floats per copy) and of the constant basis blocks (3 fp16
//                slices = 19.5 KB per k-step, L2-resident) into a 5-deep shared-memory ring;
//   warps 1, 2   MMA issuers (one elected lane each): warp 1 owns Acc0 (1 MMA per k-step), warp 2 owns
//                Acc1 (5 MMAs per k-step, operands of the next k-step awaited before the last one is
//                issued so the pipe never drains); A from TMEM, B from shared memory; warp 1 also owns
//                the TMEM allocation. Each accumulator is written by ONE thread in program order, so
//                results are deterministic;
//   warps 12..27 transform: staged waveform -> tile scale -> folded, sliced A operands written straight
//                into TMEM with tcgen05.st (thread = frame row = TMEM lane; bank-conflict-free rotated
//                reads); four warp sets take every 4th k-step through a 4-deep TMEM ring;
//   warps 4..11  epilogue (two warps per TMEM lane quarter, bins split in two): TMEM -> registers,
//                (Acc0 + Acc1)^2, sparse mel projection as straight-line code from a generated
//                compile-time table (each FFT bin feeds <= 2 adjacent triangular filters). Pass 0 parks
//                the Re part of the mel sums in shared memory, pass 1 adds the Im part, applies the tile
//                scale and log10 and stores (x + 4) / 4 directly, tracking the group maximum and the
//                tile minimum.
// TMEM: columns [0,208) Acc0, [208,416) Acc1, [416,512) the A ring (4 stages x 3 planes x 8 columns).
// Clip-edge hop-row groups (reflect padding, ragged ends) are staged by the transform warps with plain
// loads instead of TMA. logmel_floor_kernel then applies the max-8 floor, touching only the tiles
// whose minimum lies below it (digital silence, fades).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "../../include/lyricalign.h"
#include "la_common.cuh"
#include "la_mel_table.inc"

namespace la {

constexpr int kNfft = 400, kHop = 160, kMels = 80, kBins = 201;
constexpr int kTileM = 128;                 // frames per tile
constexpr int kNpad = 208;                  // 201 bins padded to a multiple of 16
constexpr int kKSteps = 13;                 // per pass: 208 folded samples / 16 (UMMA_K for f16)
constexpr int kIters = 2 * kKSteps;         // k-steps per tile: pass 0 (Re) then pass 1 (Im)
constexpr int kRawRows = 130;               // hop rows staged per tile
constexpr int kRawGroup = 8;                // hop rows per bulk copy (the TMA request rate, ~35 ns each, is the limit)
constexpr int kRawGroupPitch = kRawGroup * kHop + 8;   // floats: +8 per group -> thread-per-row reads hit 32 banks
constexpr int kRawGroups = (kRawRows + kRawGroup - 1) / kRawGroup;   // 17
constexpr int kRawBytes = kRawGroups * kRawGroupPitch * 4;           // 87584
constexpr int kBLbo = (kNpad / 8) * 128;    // 3328 bytes between K-adjacent core matrices of the basis
constexpr int kBBytes = 2 * kBLbo;          // 6656: one fp16 slice of one k-step [ki 2][ni 26][8 bins][8 samples]
constexpr int kSlices = 3;
constexpr int kBStageBytes = kSlices * kBBytes;   // 19968: [B1 B2 B3]
constexpr int kBStages = 5;
constexpr int kAStages = 4;                 // A operands live in TMEM: 3 planes x 8 columns per stage
constexpr int kAStageCols = 24;
constexpr int kACol0 = 2 * kNpad;           // TMEM columns [0,416) accumulators, [416,512) A ring
constexpr int kXformWarps = 16;             // warps 12..27: 4 sets (k-step mod 4) x 4 lane quarters
constexpr int kEpiWarps = 8;                // warps 4..11: two per TMEM lane quarter, bins split at kSplit
constexpr int kSplit = LA_MEL_SPLIT;        // bins [0, 96) -> warps 4..7, [96, 201) -> warps 8..11
constexpr int kLogmelThreads = 896;
constexpr uint32_t kTmemCols = 512;
constexpr int kPartBytes = kMels * kTileM * 4;          // Re part of the mel sums between the passes
constexpr int kSharedBytes = 2 * 2 * kTileM * 4;        // the two filters fed by both bin halves, double-buffered
constexpr int kMiscBytes = 256;                         // per-warp maxima [2][16], tile exponent [2]
static_assert(kACol0 + kAStages * kAStageCols <= 512, "TMEM columns");

struct ClipDesc {
    int64_t wave_off;   // float offset of the clip's first sample
    int64_t out_off;    // float offset of out[clip][0][0]
    int32_t n_samples;
    int32_t n_frames;
    int32_t out_stride; // floats between mel rows
    int32_t group;      // clips of one group share the max-8 floor (one whisper call)
    int32_t tile0;      // first tile of the clip
    int32_t pad;
};

struct LogmelParams {
    const float* wave;
    float* out;
    const ClipDesc* clips;
    const int32_t* tile_clip;     // [n_tiles]
    int n_tiles;
    int n_clips;
    const __half* basis;          // [pass 2][k-step 13][slice 3] canonical UMMA K-major blocks
    int* group_max;               // running maxima of the mel power (bit pattern of a float >= 0)
    int* tile_min;                // per-tile minima of the mel power (same encoding)
    int dbg;                      // LA_LOGMEL_DBG (perf triage only; 32.. break the numerics): 8 = CTA 0 timestamps its first tiles, 16 = swap the fp16 pair
                                  // order, 32 = alternate accumulators, 64 = no tcgen05.st, 128 = no epilogue math, 256 = no tile-scale scan
};

// LA_LOGMEL_DBG & 8 launches the TRACE instantiation: CTA 0 timestamps its first 8 tiles (perf triage only;
// compiled out of the production kernel)
constexpr int kTraceEvents = 64;
__device__ unsigned long long g_trace[4 * 8 * kTraceEvents];
template <bool TRACE>
__device__ __forceinline__ void trace_ev(int role, uint32_t tl, int ev) {
    if constexpr (TRACE) {
        if (blockIdx.x == 0 && tl < 8) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            g_trace[(role * 8 + tl) * kTraceEvents + ev] = t;
        }
    }
}
#define trace(dbg, role, tl, ev) trace_ev<TRACE>(role, tl, ev)

// ---- tcgen05 wrappers ---------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// A operand from TMEM (128 lanes x 8 columns = 16 fp16 along K), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The issuer warps run warp-uniform code (so ptxas keeps descriptors and addresses in uniform registers: the
// divergent `if (lane == 0)` form cost six R2UR per MMA, and the issuing thread competes for issue slots with
// six busy warps on its scheduler); one elected lane issues.
__device__ __forceinline__ void umma_f16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// Ring indexing restarts with every tile (stage = i % n for the tile's k-step i = 0..25), so a fully unrolled
// tile has compile-time stage indices. Stage s is used upt(s) = ceil((26 - s) / n) times per tile; the phase
// parity of its next use is therefore (tl * upt(s) + i / n) & 1.
__device__ __forceinline__ uint32_t ring_parity(uint32_t tl, int i, int n) {
    const int s = i % n;
    const int upt = (kIters - s + n - 1) / n;
    return (tl * (uint32_t)upt + (uint32_t)(i / n)) & 1u;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start>>4 at
// [0,14), leading (K-direction core-matrix) byte offset>>4 at [16,30), stride (M/N-direction)
// byte offset>>4 at [32,46), version 1 at [46,48), layout type 0 at [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D = f32 (1 @ [4,6)), A = B = f16 (0 @ [7,10), [10,13)), both
// K-major, N>>3 @ [17,23), M>>4 @ [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kNpad >> 3) << 17) |
                            ((uint32_t)(kTileM >> 4) << 24);

// How the hop-row groups (8 rows = 1280 samples each) of a tile are staged. Computed once per tile:
// groups [0, need) are needed at all (the rest feed only frames past the clip end), of those the
// groups [lo, hi) lie fully inside the clip and take one TMA bulk copy each, the others (clip
// edges: reflect padding, ragged end) are filled by plain loads.
struct RawPlan { int need, lo, hi; };
// 32-bit arithmetic with constant divisors only (n_samples < 2^31): every thread of the transform warps
// and the producer evaluate this once per tile, and the first version's 64-bit divisions cost 1.6 us there.
__device__ __forceinline__ RawPlan raw_plan(const ClipDesc& c, int j0, int f0) {
    constexpr int G = kRawGroup * kHop;                                  // 1280 samples per group
    RawPlan r;
    const int last_valid = min(kTileM, c.n_frames - f0) - 1;            // last frame row of the tile that is stored
    const int need_hi = last_valid * kHop + kNfft;                      // tile-relative, exclusive
    r.need = min(kRawGroups, (need_hi + G - 1) / G);
    r.lo = j0 < 0 ? 1 : 0;
    const int room = c.n_samples - j0;                                   // samples from the tile's first sample to the clip end (> 0)
    int inside = room / G;                                               // full 8-row groups ending at or before the clip end
    if (inside == kRawGroups - 1 && room >= (kRawGroups - 1) * G + (kRawRows - (kRawGroups - 1) * kRawGroup) * kHop)
        inside = kRawGroups;                                             // the last group holds only 2 rows
    r.hi = max(r.lo, min(r.need, inside));
    if ((c.wave_off & 3) != 0) r.hi = r.lo;                              // unaligned clip: no TMA at all
    return r;
}

// mel power -> (log10(max(v, 1e-10)) + 4) / 4 as ONE SFU op + ONE FFMA: log10(v) / 4 = lg2(v) * (log10(2) / 4). The accurate
// log10f is ~25 instructions; 40 of them per thread in the finish phase saturated the schedulers exactly while the
// transform warps were computing the next tile's scale (3-4 us per tile). lg2.approx is within 2^-22 absolute,
// i.e. 1e-8 of the output, four orders below the 2.5e-5 budget.
__device__ __forceinline__ float mel_to_y(float v) {
    return fmaf(__log2f(fmaxf(v, 1e-10f)), 0.07525749891599529f, 1.0f);
}
__device__ __forceinline__ float max4abs(const float4 v, float m) {
    return fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fmaxf(fabsf(v.y), fabsf(v.z)), fabsf(v.w)));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {       // element with the lower K index in the low half
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool TRACE>
__global__ void __launch_bounds__(kLogmelThreads, 1) logmel_kernel(const LogmelParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* bstages = smem;
    float* raw = reinterpret_cast<float*>(smem + kBStages * kBStageBytes);
    float* part = reinterpret_cast<float*>(smem + kBStages * kBStageBytes + kRawBytes);          // [80][128]
    float* shared2 = part + kMels * kTileM;                                                       // [2][2][128]
    float* wmax = shared2 + 2 * 2 * kTileM;                                                       // [2][16]
    int* tile_e = reinterpret_cast<int*>(wmax + 32);                                              // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBStages * kBStageBytes + kRawBytes + kPartBytes +
                                                 kSharedBytes + kMiscBytes);
    uint64_t* b_full = bars;                         // [kBStages] basis block landed (tx)
    uint64_t* b_empty = b_full + kBStages;           // [kBStages] MMAs that read it retired
    uint64_t* a_full = b_empty + kBStages;           // [kAStages] 4 transform warps stored their planes
    uint64_t* a_empty = a_full + kAStages;           // [kAStages] MMAs that read it retired
    uint64_t* raw_full = a_empty + kAStages;
    uint64_t* raw_empty = raw_full + 1;
    uint64_t* tmem_full = raw_full + 2;
    uint64_t* tmem_empty = raw_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_full + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2); }      // 2 MMA issuers
        for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 2); }
        mbar_init(raw_full, 1);
        mbar_init(raw_empty, kXformWarps);
        mbar_init(tmem_full, 2);
        mbar_init(tmem_empty, kEpiWarps);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    for (int i = tid; i < kRawBytes / 16; i += kLogmelThreads)        // pads / last group's tail must read as 0.0 in the scale scan
        reinterpret_cast<float4*>(raw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();                                               // generic-proxy zeroes before the async-proxy (TMA) writes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // all 512 columns are ours (one CTA per SM), so the allocation can only start at lane 0, column 0; using the
    // literal lets every TMEM address below be an immediate
    constexpr uint32_t tmem_base = 0;
    if (*tmem_slot != 0u) __trap();

    if (warp == 0) {
        // =============================== producer ==========================================
        if (lane == 0) {
            uint32_t tl = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
                const ClipDesc c = p.clips[p.tile_clip[tile]];
                const int f0 = (tile - c.tile0) * kTileM;
                const int j0 = f0 * kHop - kNfft / 2;
                mbar_wait(raw_empty, (tl & 1) ^ 1);
                trace(p.dbg, 3, tl, 0);
                const RawPlan rp = raw_plan(c, j0, f0);
                uint32_t tx = 0;
                for (int g = rp.lo; g < rp.hi; ++g) tx += min(kRawGroup, kRawRows - g * kRawGroup) * kHop * 4;
                if (tx) mbar_arrive_expect_tx(raw_full, tx);
                else mbar_arrive(raw_full);
                const float* src = p.wave + c.wave_off + j0;
                for (int g = rp.lo; g < rp.hi; ++g)
                    bulk_g2s(raw + g * kRawGroupPitch, src + g * kRawGroup * kHop,
                             min(kRawGroup, kRawRows - g * kRawGroup) * kHop * 4, raw_full);
                // the staging buffer is single, so the next tile's waveform can only be COPIED once this
                // tile is transformed -- but it can already be pulled into L2
                const int nt = tile + gridDim.x;
                if (nt < p.n_tiles) {
                    const ClipDesc cn = p.clips[p.tile_clip[nt]];
                    const int fn = (nt - cn.tile0) * kTileM;
                    const int jn = fn * kHop - kNfft / 2;
                    const RawPlan rn = raw_plan(cn, jn, fn);
                    if (rn.hi > rn.lo)
                        bulk_prefetch_l2(p.wave + cn.wave_off + jn + rn.lo * kRawGroup * kHop,
                                         (uint32_t)(min(rn.hi * kRawGroup, kRawRows) - rn.lo * kRawGroup) * kHop * 4);
                }
                trace(p.dbg, 3, tl, 1);
                for (int i = 0; i < kIters; ++i) {
                    const int s = i % kBStages;
                    mbar_wait(&b_empty[s], ring_parity(tl, i, kBStages) ^ 1);
                    trace(p.dbg, 3, tl, 2 + i);
                    mbar_arrive_expect_tx(&b_full[s], kBStageBytes);
                    bulk_g2s(bstages + s * kBStageBytes,
                             reinterpret_cast<const unsigned char*>(p.basis) + (size_t)i * kBStageBytes, kBStageBytes,
                             &b_full[s]);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =============================== MMA issuers =======================================
        // warp 1: Acc0 += A1*B1 (the exact chain); warp 2: Acc1 += the five cross terms. Whole warps, uniform
        // control flow, one elected lane issues; the tile's 26 k-steps are fully unrolled so every stage
        // index, TMEM column and descriptor offset is an immediate. warp 2 waits for the NEXT k-step's
        // operands between its 4th and 5th MMA, so its barrier bookkeeping overlaps MMAs still queued.
        const bool cross = warp == 2;
        const uint32_t d_acc = tmem_base + (cross ? kNpad : 0);
        const uint32_t ta0 = tmem_base + kACol0;
        const uint64_t bd0 = umma_desc(smem_u32(bstages), kBLbo, 128);
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                const int pass = i / kKSteps, ks = i % kKSteps;
                const int sb = i % kBStages, sa = i % kAStages;
                if (ks == 0) {
                    mbar_wait(tmem_empty, ((2 * tl + pass) & 1) ^ 1);   // epilogue of the previous pass drained TMEM
                    tc_fence_after();
                    if (cross && lane == 0) trace(p.dbg, 0, tl, pass ? 29 : 0);
                }
                // warp 2 has already awaited these operands (look-ahead below), except for the tile's first k-step.
                // (No look-ahead across tiles: k-step 25 and the next tile's k-step 0 share basis stage 0, whose
                // refill waits for THIS k-step's commits -- waiting for it before committing would deadlock.)
                if (!cross || i == 0) {
                    mbar_wait(&b_full[sb], ring_parity(tl, i, kBStages));
                    mbar_wait(&a_full[sa], ring_parity(tl, i, kAStages));
                    tc_fence_after();
                }
                const uint32_t a1 = ta0 + sa * kAStageCols, a2 = a1 + 8, a3 = a1 + 16;
                const uint64_t b1 = bd0 + (uint64_t)((sb * kBStageBytes) >> 4);          // start-address field += stage
                const uint64_t b2 = b1 + (uint64_t)(kBBytes >> 4), b3 = b1 + (uint64_t)((2 * kBBytes) >> 4);
                if (!cross) {
                    umma_f16_ts_elect(d_acc, a1, b1, kIdesc, ks > 0 ? 1u : 0u);
                } else {
                    if (TRACE && i >= 4 && i < 8 && lane == 0) trace(p.dbg, 0, tl, 32 + (i - 4) * 5);
                    umma_f16_ts_elect(d_acc, a3, b1, kIdesc, ks > 0 ? 1u : 0u);   // small terms first
                    umma_f16_ts_elect(d_acc, a1, b3, kIdesc, 1u);
                    umma_f16_ts_elect(d_acc, a2, b2, kIdesc, 1u);
                    umma_f16_ts_elect(d_acc, a2, b1, kIdesc, 1u);
                    if (TRACE && i >= 4 && i < 8 && lane == 0) trace(p.dbg, 0, tl, 33 + (i - 4) * 5);
                    if (i + 1 < kIters) {
                        mbar_wait(&b_full[(i + 1) % kBStages], ring_parity(tl, i + 1, kBStages));
                        mbar_wait(&a_full[(i + 1) % kAStages], ring_parity(tl, i + 1, kAStages));
                        tc_fence_after();
                    }
                    if (TRACE && i >= 4 && i < 8 && lane == 0) trace(p.dbg, 0, tl, 34 + (i - 4) * 5);
                    umma_f16_ts_elect(d_acc, a1, b2, kIdesc, 1u);
                    if (TRACE && i >= 4 && i < 8 && lane == 0) trace(p.dbg, 0, tl, 35 + (i - 4) * 5);
                }
                umma_commit_elect(&a_empty[sa]);
                umma_commit_elect(&b_empty[sb]);
                if (TRACE && cross && i >= 4 && i < 8 && lane == 0) trace(p.dbg, 0, tl, 36 + (i - 4) * 5);
                if (ks == kKSteps - 1) {
                    umma_commit_elect(tmem_full);
                    if (cross && lane == 0) trace(p.dbg, 0, tl, 27 + pass);
                }
            }
        }
    } else if (warp >= 12) {
        // ====================== transform: staged waveform -> A operands in TMEM =============
        // Thread = frame row (TMEM lane). Warp set s = (warp - 12) / 4 serves the k-steps with
        // (global k-step index) % 4 == s, i.e. always TMEM stage s. Sample j of the tile sits at
        // raw[j + 8 * (j / 1280)] (8 hop rows per bulk copy, 8 floats of padding between copies). Lane l
        // reads each 8-sample window rotated by l & 7, which makes every LDS hit 32 distinct banks;
        // three select stages undo the rotation.
        const int set = (warp - 12) >> 2;
        const int q = warp & 3;                    // TMEM lane quarter == warp % 4
        const int r = q * 32 + lane;
        const int xt = tid - 384;
        const int rot = lane & 7;
        int jj[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) jj[j] = (j + rot) & 7;
        const float* row_d0 = raw + r * kHop + 8 * (r >> 3);          // hop-row offset 0, 1, 2 of frame r
        const float* row_d1 = raw + r * kHop + 8 * ((r + 1) >> 3);
        const float* row_d2 = raw + r * kHop + 8 * ((r + 2) >> 3);
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + kACol0 + set * kAStageCols;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
            const ClipDesc c = p.clips[p.tile_clip[tile]];
            const int f0 = (tile - c.tile0) * kTileM;
            const int j0 = f0 * kHop - kNfft / 2;
            const RawPlan rp = raw_plan(c, j0, f0);                // before the wait: off the critical path
            mbar_wait(raw_full, tl & 1);
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 0);
            // ---- tile scale: largest |sample| over everything a stored frame of this tile reads ----
            float m = 0.f;
            if (rp.lo > 0 || rp.hi < rp.need) {
                // clip edges: reflect padding (torch.stft centre=True) / ragged ends, by plain loads
                const float* x = p.wave + c.wave_off;
                const int N = c.n_samples;
                for (int g = 0; g < rp.need; ++g) {
                    if (g >= rp.lo && g < rp.hi) continue;
                    const int cnt = min(kRawGroup, kRawRows - g * kRawGroup) * kHop;
                    float v[3];                                     // <= 1280 samples per group: all three loads in flight at once
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int i = xt + 512 * u;
                        int j = j0 + g * kRawGroup * kHop + i;
                        if (j < 0) j = -j;
                        if (j >= N) j = 2 * (N - 1) - j;
                        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
                        v[u] = i < cnt ? __ldg(x + j) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int i = xt + 512 * u;
                        if (i < cnt) raw[g * kRawGroupPitch + i] = v[u];
                        m = fmaxf(m, fabsf(v[u]));
                    }
                }
            }
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 32);
            {
                // the TMA-staged groups [lo, hi) are one contiguous range of the staging buffer (the 8-float pads
                // between groups and the unused tail of the last group were zeroed once at kernel start and are
                // never written), so the scan is a flat float4 loop: ~11 loads per thread
                const float4* src = reinterpret_cast<const float4*>(raw);
                const int i1 = ((p.dbg & 256) ? rp.lo : rp.hi) * (kRawGroupPitch / 4);
                for (int i = rp.lo * (kRawGroupPitch / 4) + xt; i < i1; i += 512) m = max4abs(src[i], m);
            }
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 33);
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) wmax[(tl & 1) * 16 + (warp - 12)] = m;
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 31);
            named_bar_sync(2, 512);
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 29);                                 // staging complete, per-warp maxima visible
            m = wmax[(tl & 1) * 16 + (lane & 15)];
#pragma unroll
            for (int o = 8; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            // S = 2^(ex - 125) > 2 max|x| >= |fold|; T = fold * 2^14 / S; X = acc * S * 2^-28
            int ex = (int)(__float_as_uint(m) >> 23);
            if (p.dbg & 256) ex = 129;
            ex = min(max(ex, 27), 250);
            const float scale_t = __uint_as_float((uint32_t)(266 - ex) << 23);
            if (xt == 0) tile_e[tl & 1] = ex;
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 1);
            for (int local = set; local < kIters; local += kAStages) {
                const int pass = local >= kKSteps ? 1 : 0;
                const int ks = local - pass * kKSteps;
                uint32_t pk[24];                                     // [plane 3][column 8], two fp16 per column
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n0 = ks * 16 + h * 8;                  // n = n0 + jj in 0..207
                    const float* fwd_row = (n0 >= kHop ? row_d1 : row_d0) + n0;   // 8-aligned windows never straddle 160
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int mrev = kNfft - n0 - jj[j];         // 193..400
                        const float fwd = fwd_row[jj[j]];
                        const float rev = (mrev >= 2 * kHop ? row_d2 : row_d1)[mrev];
                        v[j] = pass ? fwd - rev : fwd + rev;          // n = 200: e = 2 x[200] (basis row halved), o = 0
                    }
                    // undo the rotation: t[c] = v[(c - rot) & 7]
                    float t[8];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) t[cidx] = (rot & 1) ? v[(cidx + 7) & 7] : v[cidx];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) v[cidx] = (rot & 2) ? t[(cidx + 6) & 7] : t[cidx];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) t[cidx] = (rot & 4) ? v[(cidx + 4) & 7] : v[cidx];
                    if (ks == 0 && h == 0) t[0] = 0.f;                // n = 0 pairs x[0] with x[400], which the last frame may not have staged (basis row 0 is zero)
                    // slice on the fixed grid: A1 = multiple of 64, A2 = fp16(T - A1), A3 = fp16(T - A1 - A2)
#pragma unroll
                    for (int cidx = 0; cidx < 8; cidx += 2) {
                        const float T0 = t[cidx] * scale_t, T1 = t[cidx + 1] * scale_t;
                        const float A0 = (T0 + 805306368.f) - 805306368.f;          // 1.5 * 2^29: ulp 64, round to nearest even
                        const float A1 = (T1 + 805306368.f) - 805306368.f;
                        const float R0 = T0 - A0, R1 = T1 - A1;                       // exact
                        const __half2 h2 = __floats2half2_rn(R0, R1);
                        const float2 f2 = __half22float2(h2);
                        const int col = h * 4 + (cidx >> 1);
                        if (p.dbg & 16) {
                            pk[col] = pack_h2(A1, A0);
                            pk[8 + col] = pack_h2(f2.y, f2.x);
                            pk[16 + col] = pack_h2(R1 - f2.y, R0 - f2.x);
                        } else {
                            pk[col] = pack_h2(A0, A1);
                            pk[8 + col] = *reinterpret_cast<const uint32_t*>(&h2);
                            pk[16 + col] = pack_h2(R0 - f2.x, R1 - f2.y);
                        }
                    }
                }
                // only the stores need the TMEM stage: everything above overlaps the MMAs in flight
                if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 16 + (local >> 2));
                mbar_wait(&a_empty[set], ring_parity(tl, local, kAStages) ^ 1);
                tc_fence_after();
                if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 2 + (local >> 2));
                if (!(p.dbg & 64)) {
                    uint32_t pl[8];
#pragma unroll
                    for (int pln = 0; pln < 3; ++pln) {
#pragma unroll
                        for (int cidx = 0; cidx < 8; ++cidx) pl[cidx] = pk[pln * 8 + cidx];
                        tmem_st8(tlane + pln * 8, pl);
                    }
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[set]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(raw_empty);      // staged waveform no longer needed
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 30);
        }
    } else if (warp >= 4 && warp < 12) {
        // ====================== epilogue: (Acc0 + Acc1)^2 -> mel, straight out of TMEM (256 threads) ====
        // Two warps share each TMEM lane quarter and split the 201 bins at kSplit. Exactly two filters
        // (ms, ms + 1, ms = first filter fed by bin kSplit) receive power from both sides: the high half
        // parks its share in `shared2`, the low half finishes and stores them after one named barrier.
        // The loop body is kept small on purpose: the first version unrolled to 3200 instructions and
        // spent its time in instruction fetch.
        const int half = warp >= 8 ? 1 : 0;
        const int wq = warp & 3;                   // TMEM lane quarter == warp % 4
        const int row = wq * 32 + lane;
        constexpr int ms = LA_MEL_MS;
        float r_ms0 = 0.f, r_ms1 = 0.f;            // low half: Re part of the two shared filters
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
            const ClipDesc c = p.clips[p.tile_clip[tile]];
            const int f0 = (tile - c.tile0) * kTileM;
            const int f = f0 + row;
            const bool valid = f < c.n_frames;
            float* sh2 = shared2 + (tl & 1) * 2 * kTileM + row;
            for (int pass = 0; pass < 2; ++pass) {
                const uint32_t ph = 2 * tl + pass;
                mbar_wait(tmem_full, ph & 1);
                tc_fence_after();
                if (warp == 4 && lane == 0) trace(p.dbg, 2, tl, 2 * pass);
                // The drain is on the tensor pipe's critical path (the next pass cannot start before TMEM is
                // read out), so it only accumulates: pass 0 parks the Re part of every mel sum in `part`,
                // pass 1 adds the Im part. log10 and the stores happen after TMEM is released.
                float* sp = part + (half ? ms : 0) * kTileM + row;
                float a0 = 0.f, a1 = 0.f;
                auto flush = [&]() {                    // the filter held in a0 is complete for this pass
                    *sp = pass == 0 ? a0 : *sp + a0;
                    sp += kTileM;
                    a0 = a1; a1 = 0.f;
                };
                int nshared = 0;
                auto flush_hi = [&]() {                 // high half: its first two filters are the shared ones
                    float* dst = nshared < 2 ? sh2 + nshared * kTileM : sp;
                    *dst = pass == 0 ? a0 : *dst + a0;
                    ++nshared;
                    sp += kTileM;
                    a0 = a1; a1 = 0.f;
                };
                const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
                uint32_t c0[16], c1[16];
                auto release_tmem = [&]() {             // this warp's share of TMEM is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty);
                };
                // Straight-line code: the filterbank is a compile-time table (la_mel_table.inc), so the
                // weights are FFMA immediates and the flush points are static. The first bin of each
                // half never flushes (the accumulators start empty on that filter).
#define LA_BIN(K, LO, W0, W1, NF)                                                                 \
                {                                                                                     \
                    const float x = __uint_as_float(c0[(K) & 15]) + __uint_as_float(c1[(K) & 15]);    \
                    const float pw = x * x;                                                           \
                    if ((NF) >= 1 && (K) != 0 && (K) != kSplit) LA_FLUSH();                           \
                    if ((NF) >= 2 && (K) != 0 && (K) != kSplit) LA_FLUSH();                           \
                    a0 = fmaf(W0, pw, a0);                                                            \
                    a1 = fmaf(W1, pw, a1);                                                            \
                }
#define LA_CHUNK(C, LAST)                                                                         \
                {                                                                                     \
                    tmem_ld16(lane_base + 16 * (C), c0);                                              \
                    tmem_ld16(lane_base + kNpad + 16 * (C), c1);                                      \
                    tmem_ld_wait();                                                                   \
                    if (LAST) release_tmem();                                                         \
                    LA_MEL_CHUNK_##C(LA_BIN)                                                          \
                }
                if (p.dbg & 128) {
                    release_tmem();
                } else if (!half) {
#define LA_FLUSH flush
                    LA_CHUNK(0, false) LA_CHUNK(1, false) LA_CHUNK(2, false)
                    LA_CHUNK(3, false) LA_CHUNK(4, false) LA_CHUNK(5, true)
#pragma unroll
                    for (int i = 0; i < LA_MEL_TAIL0; ++i) flush();
#undef LA_FLUSH
                } else {
#define LA_FLUSH flush_hi
                    LA_CHUNK(6, false) LA_CHUNK(7, false) LA_CHUNK(8, false) LA_CHUNK(9, false)
                    LA_CHUNK(10, false) LA_CHUNK(11, false) LA_CHUNK(12, true)
#pragma unroll
                    for (int i = 0; i < LA_MEL_TAIL1; ++i) flush_hi();
#undef LA_FLUSH
                }
#undef LA_CHUNK
#undef LA_BIN
                if (warp == 4 && lane == 0) trace(p.dbg, 2, tl, 2 * pass + 1);
                if (pass == 0) {
                    if (!half) { r_ms0 = a0; r_ms1 = a1; }      // the low half's Re share of filters ms, ms + 1
                    continue;
                }
                // ---- finish (off the critical path: the next tile's MMAs are already running) ----
                named_bar_sync(3, 256);                         // the high half's shares of filters ms, ms+1 are in sh2
                if (!half) {
                    part[ms * kTileM + row] = sh2[0] + r_ms0 + a0;
                    part[(ms + 1) * kTileM + row] = sh2[kTileM] + r_ms1 + a1;
                }
                named_bar_sync(3, 256);                         // `part` holds the unscaled mel power of the whole tile
                const float sc = __uint_as_float((uint32_t)(tile_e[tl & 1] - 26) << 23);    // S * 2^-28
                const float sc2 = sc * sc;
                float mx = 0.f, mn = INFINITY;
                if (valid) {
                    const float* src = part + half * (kMels / 2) * kTileM + row;
                    float* dst = p.out + c.out_off + f + (int64_t)half * (kMels / 2) * c.out_stride;
#pragma unroll 8
                    for (int m = 0; m < kMels / 2; ++m, src += kTileM, dst += c.out_stride) {
                        const float v = *src * sc2;
                        *dst = mel_to_y(v);
                        mx = fmaxf(mx, v);
                        mn = fminf(mn, v);
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                }
                if (lane == 0) {                                // powers are >= 0: int order == float order
                    atomicMax(p.group_max + c.group, __float_as_int(mx));
                    atomicMin(p.tile_min + tile, __float_as_int(mn));
                }
                named_bar_sync(3, 256);                         // `part` is free for the next tile's pass 0
                if (warp == 4 && lane == 0) trace(p.dbg, 2, tl, 4);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// max(x, gmax - 8) on the stored (x + 4) / 4 values; h(z) = (z + 4) / 4 is monotone, so
// h(max(x, g)) == max(h(x), h(g)) bit for bit. One block per tile; tiles whose smallest mel power is
// above the floor (almost all of them: the floor is 80 dB below the loudest cell) exit at once.
__global__ void logmel_floor_kernel(const LogmelParams p) {
    const int tile = blockIdx.x;
    const ClipDesc c = p.clips[p.tile_clip[tile]];
    const float floor_y = mel_to_y(__int_as_float(p.group_max[c.group])) - 2.0f;    // (x - 8 + 4) / 4 == (x + 4) / 4 - 2
    const float min_y = mel_to_y(__int_as_float(p.tile_min[tile]));
    if (min_y >= floor_y) return;
    const int f = (tile - c.tile0) * kTileM + threadIdx.x;     // one frame column per thread, coalesced along f
    if (f >= c.n_frames) return;
    float* q = p.out + c.out_off + f;
#pragma unroll 8
    for (int m = 0; m < kMels; ++m, q += c.out_stride)
        if (*q < floor_y) *q = floor_y;
}

__global__ void logmel_init_kernel(int* group_max, int n_groups, int* tile_min, int n_tiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups) group_max[i] = 0;                        // 0.0f: powers are >= 0
    if (i < n_tiles) tile_min[i] = 0x7f800000;                 // +inf
}

// ---- host: constant tables ---------------------------------------------------------------
static uint16_t f32_to_f16_rn(float f) {       // IEEE round-to-nearest-even, subnormals kept
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x7fffffu;
    const int exp = (int)((x >> 23) & 0xffu);
    if (exp == 255) return (uint16_t)(sign | 0x7c00u | (mant ? 0x200u : 0u));
    const int e = exp - 127 + 15;
    if (e >= 31) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        mant |= 0x800000u;
        const int shift = 14 - e;                                   // 14..24
        uint32_t h = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)e << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;         // a carry into the exponent is the right answer
    return (uint16_t)(sign | h);
}
static double f16_to_f64(uint16_t h) {
    const int e = (h >> 10) & 0x1f, m = h & 0x3ff;
    double v;
    if (e == 0) v = std::ldexp((double)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = std::ldexp((double)(m | 0x400), e - 25);
    return (h & 0x8000) ? -v : v;
}

// basis blocks [pass 2][k-step 13][slice 3], each [ki 2][ni 26][8 rows (bin)][8 fp16 (sample)] -- the UMMA
// canonical K-major no-swizzle layout. Slices of U = 2^14 w[n] cos|(-sin)(2 pi k n / 400) (fp64):
// B1 = 64 rint(U / 64), B2 = fp16(U - B1), B3 = fp16(U - B1 - B2).
static void build_basis(std::vector<uint16_t>& h) {
    h.assign((size_t)kIters * kSlices * (kBBytes / 2), 0);
    const double PI = 3.14159265358979323846;
    for (int pass = 0; pass < 2; ++pass)
        for (int ks = 0; ks < kKSteps; ++ks)
            for (int kk = 0; kk < 16; ++kk) {
                const int n = ks * 16 + kk;                                   // 0..207; rows 0 and 201.. are zero
                if (n < 1 || n > kNfft / 2) continue;
                const double w = 0.5 - 0.5 * std::cos(2.0 * PI * n / kNfft);
                for (int b = 0; b < kBins; ++b) {
                    const int ph = (int)(((long long)b * n) % kNfft);         // exact phase reduction
                    double v;
                    // n = 200 folds onto itself: the kernel forms e[200] = 2 x[200], so halve its row
                    if (pass == 0) v = (n == kNfft / 2 ? 0.5 : 1.0) * w * std::cos(2.0 * PI * ph / kNfft);
                    else v = (n == kNfft / 2) ? 0.0 : -w * std::sin(2.0 * PI * ph / kNfft);
                    const double U = v * 16384.0;
                    const double B1 = 64.0 * std::nearbyint(U / 64.0);
                    const uint16_t h1 = f32_to_f16_rn((float)B1);
                    const uint16_t h2 = f32_to_f16_rn((float)(U - B1));
                    const uint16_t h3 = f32_to_f16_rn((float)(U - B1 - f16_to_f64(h2)));
                    const size_t inner = (size_t)(kk >> 3) * (kBLbo / 2) + (size_t)(b >> 3) * 64 + (b & 7) * 8 + (kk & 7);
                    const size_t base = (size_t)(pass * kKSteps + ks) * kSlices * (kBBytes / 2);
                    h[base + inner] = h1;
                    h[base + (kBBytes / 2) + inner] = h2;
                    h[base + 2 * (kBBytes / 2) + inner] = h3;
                }
            }
}

struct LogmelTables {
    std::mutex mu;
    __half* d_basis[64] = {nullptr};
};
static LogmelTables g_tab;

static cudaError_t ensure_tables(int device, const __half** basis_out) {
    std::lock_guard<std::mutex> lock(g_tab.mu);
    if (!g_tab.d_basis[device]) {
        std::vector<uint16_t> h;
        build_basis(h);
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, h.size() * 2);
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(d); return e; }
        g_tab.d_basis[device] = static_cast<__half*>(d);
    }
    *basis_out = g_tab.d_basis[device];
    return cudaSuccess;
}
// la_shutdown(): the per-device basis tables are the only device memory K1 keeps between calls
void logmel_release_tables() {
    std::lock_guard<std::mutex> lock(g_tab.mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 64; ++d)
        if (g_tab.d_basis[d]) {
            cudaSetDevice(d);
            cudaFree(g_tab.d_basis[d]);
            g_tab.d_basis[d] = nullptr;
        }
    cudaSetDevice(cur);
}

size_t logmel_smem_bytes() {
    return (size_t)kBStages * kBStageBytes + kRawBytes + kPartBytes + kSharedBytes + kMiscBytes + 256;
}

}  // namespace la

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static thread_local char g_lm_err[256];   // scratch for formatting; the message is handed to la::set_error
#define LM_FAIL(code) return la::set_error(code, g_lm_err)

static inline size_t lm_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int logmel_run_impl(const float* d_wave, float* d_out, const std::vector<la::ClipDesc>& clips_in, int n_groups,
                           void* d_ws, size_t ws_bytes, cudaStream_t stream);

static int logmel_run(const float* d_wave, float* d_out, const std::vector<la::ClipDesc>& clips_in, int n_groups,
                      void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    try {                                    // nothing may throw across the C ABI
        return logmel_run_impl(d_wave, d_out, clips_in, n_groups, d_ws, ws_bytes, stream);
    } catch (...) {
        snprintf(g_lm_err, sizeof g_lm_err, "out of host memory");
        LM_FAIL(LA_ERR_ALLOC);
    }
}

static int logmel_run_impl(const float* d_wave, float* d_out, const std::vector<la::ClipDesc>& clips_in, int n_groups,
                           void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    using namespace la;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "cudaGetDevice: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    const __half* basis = nullptr;
    e = ensure_tables(device, &basis);
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "tables: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    std::vector<ClipDesc> clips = clips_in;
    std::vector<int32_t> tile_clip;
    for (size_t ci = 0; ci < clips.size(); ++ci) {
        clips[ci].tile0 = (int32_t)tile_clip.size();
        const int nt = (clips[ci].n_frames + kTileM - 1) / kTileM;
        for (int t = 0; t < nt; ++t) tile_clip.push_back((int32_t)ci);
    }
    const int n_tiles = (int)tile_clip.size();
    if (n_tiles == 0) return LA_OK;
    const size_t o_clips = 0;
    const size_t o_tiles = lm_align(o_clips + clips.size() * sizeof(ClipDesc), 256);
    const size_t o_max = lm_align(o_tiles + tile_clip.size() * 4, 256);
    const size_t o_min = o_max + lm_align((size_t)n_groups * 4, 256);
    const size_t need = o_min + lm_align((size_t)n_tiles * 4, 256);
    if (ws_bytes < need) { snprintf(g_lm_err, sizeof g_lm_err, "workspace too small"); LM_FAIL(LA_ERR_ARG); }
    unsigned char* ws = static_cast<unsigned char*>(d_ws);
    e = cudaMemcpyAsync(ws + o_clips, clips.data(), clips.size() * sizeof(ClipDesc), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws + o_tiles, tile_clip.data(), tile_clip.size() * 4, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "meta upload: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    // pageable sources: the copies above complete (w.r.t. the host buffers) before returning
    LogmelParams p;
    p.wave = d_wave; p.out = d_out;
    p.clips = reinterpret_cast<const ClipDesc*>(ws + o_clips);
    p.tile_clip = reinterpret_cast<const int32_t*>(ws + o_tiles);
    p.n_tiles = n_tiles; p.n_clips = (int)clips.size();
    p.basis = basis;
    p.group_max = reinterpret_cast<int*>(ws + o_max);
    p.tile_min = reinterpret_cast<int*>(ws + o_min);
    { const char* d = getenv("LA_LOGMEL_DBG"); p.dbg = d ? atoi(d) : 0; }
    logmel_init_kernel<<<(std::max(n_groups, n_tiles) + 255) / 256, 256, 0, stream>>>(p.group_max, n_groups, p.tile_min, n_tiles);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t smem = logmel_smem_bytes();
    static bool attr_done[64] = {};               // per-function, per-device opt-in: set once, not per launch
    if (device >= 0 && device < 64 && !attr_done[device]) {
        e = cudaFuncSetAttribute(logmel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "smem attr: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
        attr_done[device] = true;
    }
    if (p.dbg & 8) logmel_kernel<true><<<std::min(n_tiles, sms), kLogmelThreads, smem, stream>>>(p);
    else logmel_kernel<false><<<std::min(n_tiles, sms), kLogmelThreads, smem, stream>>>(p);
    logmel_floor_kernel<<<n_tiles, kTileM, 0, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "launch: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    return LA_OK;
}

extern "C" {

// perf triage only (not part of the public header): copies the LA_LOGMEL_DBG&8 timestamps out
int la_debug_logmel_trace(unsigned long long* h_out) {
    return cudaMemcpyFromSymbol(h_out, la::g_trace, sizeof(unsigned long long) * 4 * 8 * la::kTraceEvents) == cudaSuccess ? 0 : -2;
}

// test hook (not in the public header): the host-built fp16 basis table, so a CPU test can check the slicing
size_t la_debug_logmel_basis(uint16_t* out, size_t cap_elems) {
    std::vector<uint16_t> h;
    la::build_basis(h);
    if (out) memcpy(out, h.data(), std::min(cap_elems, h.size()) * 2);
    return h.size();
}

size_t la_logmel_workspace_bytes(int n_clips, int64_t total_samples) {
    if (n_clips < 0 || total_samples < 0) return 0;
    const size_t tiles = (size_t)(total_samples / la::kHop) / la::kTileM + (size_t)n_clips + 1;
    return lm_align((size_t)n_clips * sizeof(la::ClipDesc), 256) + 2 * lm_align(tiles * 4, 256) +
           lm_align((size_t)std::max(n_clips, 1) * 4, 256) + 256;
}

int la_logmel(const float* d_wave, int batch, int64_t n_samples, int64_t wave_stride, float* d_out,
              int64_t out_stride, void* d_ws, void* stream) {
    try {
        if (!d_wave || !d_out || !d_ws || batch < 0) { snprintf(g_lm_err, sizeof g_lm_err, "null argument"); LM_FAIL(LA_ERR_ARG); }
        if (reinterpret_cast<uintptr_t>(d_wave) & 15) { snprintf(g_lm_err, sizeof g_lm_err, "waveform base must be 16-byte aligned"); LM_FAIL(LA_ERR_ARG); }
        if (n_samples <= la::kNfft / 2) { snprintf(g_lm_err, sizeof g_lm_err, "reflect padding needs more than 200 samples"); LM_FAIL(LA_ERR_ARG); }
        const int64_t F = n_samples / la::kHop;
        if (out_stride < F || n_samples > INT32_MAX) { snprintf(g_lm_err, sizeof g_lm_err, "bad stride/size"); LM_FAIL(LA_ERR_ARG); }
        std::vector<la::ClipDesc> clips((size_t)batch);
        for (int b = 0; b < batch; ++b) {
            clips[b].wave_off = (int64_t)b * wave_stride;
            clips[b].out_off = (int64_t)b * la::kMels * out_stride;
            clips[b].n_samples = (int32_t)n_samples;
            clips[b].n_frames = (int32_t)F;
            clips[b].out_stride = (int32_t)out_stride;
            clips[b].group = 0;                       // one call == one global maximum (whisper semantics)
            clips[b].tile0 = 0; clips[b].pad = 0;
        }
        return logmel_run(d_wave, d_out, clips, 1, d_ws, la_logmel_workspace_bytes(batch, (int64_t)batch * n_samples),
                          static_cast<cudaStream_t>(stream));
    } catch (...) {
        snprintf(g_lm_err, sizeof g_lm_err, "out of host memory");
        LM_FAIL(LA_ERR_ALLOC);
    }
}

int la_logmel_ragged(const float* d_wave, int n_clips, const int64_t* h_wave_off, const int32_t* h_n_samples,
                     float* d_out, const int64_t* h_out_off, const int32_t* h_out_stride, void* d_ws, void* stream) {
    try {
        if (!d_wave || !d_out || !d_ws || n_clips < 0 || !h_wave_off || !h_n_samples || !h_out_off || !h_out_stride) {
            snprintf(g_lm_err, sizeof g_lm_err, "null argument");
            LM_FAIL(LA_ERR_ARG);
        }
        if (reinterpret_cast<uintptr_t>(d_wave) & 15) { snprintf(g_lm_err, sizeof g_lm_err, "waveform base must be 16-byte aligned"); LM_FAIL(LA_ERR_ARG); }
        std::vector<la::ClipDesc> clips((size_t)n_clips);
        int64_t total = 0;
        for (int c = 0; c < n_clips; ++c) {
            if (h_n_samples[c] <= la::kNfft / 2) { snprintf(g_lm_err, sizeof g_lm_err, "clip %d too short for reflect padding", c); LM_FAIL(LA_ERR_ARG); }
            clips[c].wave_off = h_wave_off[c];
            clips[c].out_off = h_out_off[c];
            clips[c].n_samples = h_n_samples[c];
            clips[c].n_frames = h_n_samples[c] / la::kHop;
            clips[c].out_stride = h_out_stride[c];
            clips[c].group = c;                       // independent calls: one maximum per clip (batch size 1)
            clips[c].tile0 = 0; clips[c].pad = 0;
            if (clips[c].out_stride < clips[c].n_frames) { snprintf(g_lm_err, sizeof g_lm_err, "clip %d: out_stride < frames", c); LM_FAIL(LA_ERR_ARG); }
            total += h_n_samples[c];
        }
        return logmel_run(d_wave, d_out, clips, std::max(n_clips, 1), d_ws, la_logmel_workspace_bytes(n_clips, total),
                          static_cast<cudaStream_t>(stream));
    } catch (...) {
        snprintf(g_lm_err, sizeof g_lm_err, "out of host memory");
        LM_FAIL(LA_ERR_ALLOC);
    }
}

}  // extern "C"
