// la_logmel.cu -- K1: Whisper-style 80-bin log-mel front end as a framed, windowed DFT on the
// 5th-gen tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, fp32 accumulators in TMEM).
//
// Replaces whisper.audio.log_mel_spectrogram as called at module/align_model.py:84 (the
// reference runs it on the CPU through torch.stft): hann(400, periodic) window, hop 160,
// centre=True reflect padding, last frame dropped, |.|^2, 80 x 201 Slaney mel filterbank,
// log10(clamp 1e-10), max(x, GLOBAL max - 8), (x + 4) / 4.
//
// Formulation. The Hann window is symmetric (w[n] = w[400-n], w[0] = 0), so with the folded
// inputs  e[n] = x[n] + x[400-n],  o[n] = x[n] - x[400-n]  (n = 1..199; e[200] = x[200]):
//     Re X[k] = sum_{n=1..200} e[n] * w[n] cos(2 pi k n / 400)
//     Im X[k] = sum_{n=1..199} o[n] * (-w[n] sin(2 pi k n / 400))
// i.e. two GEMMs [128 frames x 200] x [200 x 208] per tile -- half the FLOPs and half the basis
// bytes of the plain [128 x 400] x [400 x 402] product. Precision: every operand is split
// x = hi + lo with hi exactly representable in TF32; the kernel issues hi*hi + hi*lo + lo*hi
// (the dropped lo*lo term is 2^-22 relative), which lands 1e-6..1e-5 from the fp64 oracle in
// the log10 domain (single-pass TF32 would be 3e-1 off on tonal input).
//
// One persistent CTA per SM (896 threads), 128-frame tiles, warp-specialised:
//   warp 0       producer: 1-D TMA bulk copies of the tile's waveform (17 copies of 8 hop rows into a
//                staging area padded by 8 floats per copy) and of the constant basis blocks (26 KB per
//                k-step, L2-resident) into a 5-deep shared-memory ring;
//   warps 1, 2   MMA issuers (one elected lane each: Re and Im accumulators, 75 tcgen05.mma per tile
//                each, A operand from TMEM, B from shared memory); warp 1 also owns the TMEM allocation;
//   warps 12..27 transform: staged waveform -> folded, hi/lo-split A operands written straight into
//                TMEM with tcgen05.st (thread = frame row = TMEM lane; bank-conflict-free rotated
//                reads); two warp sets ping-pong the k-steps through a 3-deep TMEM ring;
//   warps 4..11  epilogue (two warps per TMEM lane quarter, bins split in two): TMEM -> registers,
//                power, sparse mel projection as straight-line code from a generated compile-time
//                table (each FFT bin feeds <= 2 adjacent triangular filters), coalesced stores of the
//                mel power, running max.
// TMEM: columns [0,416) hold the Re/Im accumulators, [416,512) the A ring. Clip-edge hop-row groups
// (reflect padding, ragged ends) are staged by the transform warps with plain loads instead of
// TMA. A second small kernel applies log10, the max-8 floor and (x+4)/4.
// History of the tuning (profiles/k1_tuning_r1.md): SS-mode operands were shared-memory-bandwidth
// bound, a 3200-instruction unrolled epilogue was instruction-fetch bound, a single MMA issuer left
// the tensor pipe idle ~30 % of the time.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/lyricalign.h"
#include "la_common.cuh"
#include "la_mel_table.inc"

namespace la {

constexpr int kNfft = 400, kHop = 160, kMels = 80, kBins = 201;
constexpr int kTileM = 128;                 // frames per tile
constexpr int kNpad = 208;                  // 201 bins padded to a multiple of 16
constexpr int kKSteps = 25;                 // 200 folded samples / 8 (UMMA_K for tf32)
constexpr int kRawRows = 130;               // hop rows staged per tile
constexpr int kRawGroup = 8;                // hop rows per bulk copy (the TMA request rate, ~35 ns each, is the limit)
constexpr int kRawGroupPitch = kRawGroup * kHop + 8;   // floats: +8 per group -> thread-per-row reads hit 32 banks
constexpr int kRawGroups = (kRawRows + kRawGroup - 1) / kRawGroup;   // 17
constexpr int kRawBytes = kRawGroups * kRawGroupPitch * 4;           // 87584
constexpr int kBLbo = (kNpad / 8) * 128;    // 3328 bytes between K-adjacent core matrices of the basis
constexpr int kBBytes = 2 * kBLbo;          // 6656: one basis operand (hi or lo) of one k-step
constexpr int kBStageBytes = 4 * kBBytes;   // 26624: [C_hi C_lo S_hi S_lo]
constexpr int kBStages = 5;
constexpr int kAStages = 3;                 // A operands live in TMEM: 4 planes x 8 columns per stage
constexpr int kACol0 = 2 * kNpad;           // TMEM columns [0,416) accumulators, [416,512) A ring
constexpr int kXformWarps = 16;             // warps 12..27: {even,odd k-steps} x {even,odd planes} x 4 lane quarters
constexpr int kEpiWarps = 8;                // warps 4..11: two per TMEM lane quarter, bins split at kSplit
constexpr int kSplit = LA_MEL_SPLIT;        // bins [0, 96) -> warps 4..7, [96, 201) -> warps 8..11
constexpr int kLogmelThreads = 896;
constexpr uint32_t kTmemCols = 512;
static_assert(kACol0 + kAStages * 32 <= 512, "TMEM columns");

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
    const float* basis;           // [25][even: C_hi, C_lo | odd: S_hi, S_lo] canonical UMMA layout
    int* group_max;               // running maxima of the mel power (bit pattern of a float >= 0)
    int dbg;                      // LA_LOGMEL_DBG bisect knobs (perf triage only): 1 skip epilogue math, 2 skip transform math, 4 skip basis loads
};

// LA_LOGMEL_DBG & 8: CTA 0 timestamps its first 8 tiles (perf triage only)
__device__ unsigned long long g_trace[4 * 8 * 32];
__device__ __forceinline__ void trace(int dbg, int role, uint32_t tl, int ev) {
    if ((dbg & 8) && blockIdx.x == 0 && tl < 8) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[(role * 8 + tl) * 32 + ev] = t;
    }
}



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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from TMEM (128 lanes x 8 columns of tf32), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
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
// cute::UMMA::InstrDescriptor: D = f32 (1 @ [4,6)), A = B = tf32 (2 @ [7,10), [10,13)), both
// K-major, N>>3 @ [17,23), M>>4 @ [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNpad >> 3) << 17) |
                            ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ int float_to_ordered(float v) {
    const int i = __float_as_int(v);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int k) {
    return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

__device__ __forceinline__ float rna_tf32(float x) {      // round-to-nearest TF32 (10 explicit mantissa bits)
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

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

__global__ void __launch_bounds__(kLogmelThreads, 1) logmel_kernel(const LogmelParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* bstages = smem;
    float* raw = reinterpret_cast<float*>(smem + kBStages * kBStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBStages * kBStageBytes + kRawBytes);
    uint64_t* b_full = bars;                         // [kBStages] basis block landed (tx)
    uint64_t* b_empty = b_full + kBStages;           // [kBStages] MMAs that read it retired
    uint64_t* a_full = b_empty + kBStages;           // [kAStages] 8 transform warps stored their planes
    uint64_t* a_empty = a_full + kAStages;           // [kAStages] MMAs that read it retired
    uint64_t* raw_full = a_empty + kAStages;
    uint64_t* raw_empty = raw_full + 1;
    uint64_t* tmem_full = raw_full + 2;
    uint64_t* tmem_empty = raw_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_full + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 2); }      // 2 MMA issuers
        for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], kXformWarps / 2); mbar_init(&a_empty[s], 2); }
        mbar_init(raw_full, 1);
        mbar_init(raw_empty, kXformWarps);
        mbar_init(tmem_full, 2);
        mbar_init(tmem_empty, kEpiWarps);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== producer ==========================================
        if (lane == 0) {
            uint32_t it = 0, tl = 0;
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
                for (int ks = 0; ks < kKSteps; ++ks, ++it) {
                    const int s = it % kBStages;
                    mbar_wait(&b_empty[s], ((it / kBStages) & 1) ^ 1);
                    trace(p.dbg, 3, tl, 2 + ks);
                    if (p.dbg & 4) { mbar_arrive(&b_full[s]); continue; }
                    mbar_arrive_expect_tx(&b_full[s], kBStageBytes);
                    bulk_g2s(bstages + s * kBStageBytes,
                             reinterpret_cast<const unsigned char*>(p.basis) + (size_t)ks * kBStageBytes, kBStageBytes,
                             &b_full[s]);
                }
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =============================== MMA issuers =======================================
        // Two issuing threads: warp 1 accumulates Re (even planes x cos basis), warp 2 Im (odd planes x
        // sin basis). tcgen05.mma issue is back-pressured by the tensor pipe, so a single issuer's
        // per-k-step bookkeeping (barrier waits, commits) left the pipe idle ~30 % of the time;
        // with two, one thread's bookkeeping overlaps the other's MMAs. Each accumulator is written
        // by one thread in program order, so results stay deterministic.
        if (lane == 0) {
            const int im = warp == 2 ? 1 : 0;
            const uint32_t d_acc = tmem_base + (im ? kNpad : 0);
            const uint32_t ta0 = tmem_base + kACol0 + (im ? 16 : 0);               // hi plane; lo plane at +8
            const uint64_t bd_hi0 = umma_desc(smem_u32(bstages) + (im ? 2 * kBBytes : 0), kBLbo, 128);
            const uint64_t bd_lo0 = umma_desc(smem_u32(bstages) + (im ? 3 * kBBytes : kBBytes), kBLbo, 128);
            uint32_t it = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
                mbar_wait(tmem_empty, (tl & 1) ^ 1);       // epilogue of the previous tile drained TMEM
                tc_fence_after();
                if (!im) trace(p.dbg, 0, tl, 0);
                for (int ks = 0; ks < kKSteps; ++ks, ++it) {
                    const int sb = it % kBStages, sa = it % kAStages;
                    mbar_wait(&b_full[sb], (it / kBStages) & 1);
                    mbar_wait(&a_full[sa], (it / kAStages) & 1);
                    tc_fence_after();
                    if (!im) trace(p.dbg, 0, tl, 1 + ks);
                    const uint32_t ta = ta0 + sa * 32;
                    const uint64_t b_hi = bd_hi0 + (uint64_t)(sb * (kBStageBytes >> 4));   // start-address field += stage
                    const uint64_t b_lo = bd_lo0 + (uint64_t)(sb * (kBStageBytes >> 4));
                    umma_tf32_ts(d_acc, ta, b_lo, kIdesc, ks > 0 ? 1u : 0u);   // small terms first
                    umma_tf32_ts(d_acc, ta + 8, b_hi, kIdesc, 1u);
                    umma_tf32_ts(d_acc, ta, b_hi, kIdesc, 1u);
                    umma_commit(&a_empty[sa]);
                    umma_commit(&b_empty[sb]);
                }
                umma_commit(tmem_full);
                if (!im) trace(p.dbg, 0, tl, 26);
            }
        }
    } else if (warp >= 12) {
        // ====================== transform: staged waveform -> A operands in TMEM =============
        // Thread = frame row (TMEM lane). Warps 12..15 produce the even planes (x[n] + x[400-n]),
        // warps 16..19 the odd ones (x[n] - x[400-n]); both split hi/lo and tcgen05.st 8 columns
        // per plane. Sample s of the tile sits at raw[s + 8 * (s / 1280)] (8 hop rows per bulk
        // copy, 8 floats of padding between copies). Lane l reads its 8 samples rotated by l & 7,
        // which makes every LDS hit 32 distinct banks; three select stages undo the rotation.
        const bool odd = (warp - 12) & 4;          // warps 12-15, 20-23: even planes; 16-19, 24-27: odd planes
        const int kpar = warp >= 20 ? 1 : 0;       // which k-steps this warp serves (two warp sets ping-pong)
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
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + kACol0 + (odd ? 16 : 0);
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
            const ClipDesc c = p.clips[p.tile_clip[tile]];
            const int f0 = (tile - c.tile0) * kTileM;
            const int j0 = f0 * kHop - kNfft / 2;
            const RawPlan rp = raw_plan(c, j0, f0);                // before the wait: off the critical path
            mbar_wait(raw_full, tl & 1);
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 0);
            if (rp.lo > 0 || rp.hi < rp.need) {
                // clip edges: reflect padding (torch.stft centre=True) / ragged ends, by plain loads
                const float* x = p.wave + c.wave_off;
                const int N = c.n_samples;
                for (int g = 0; g < rp.need; ++g) {
                    if (g >= rp.lo && g < rp.hi) continue;
                    const int cnt = min(kRawGroup, kRawRows - g * kRawGroup) * kHop;
                    for (int i = xt; i < cnt; i += 512) {
                        int j = j0 + g * kRawGroup * kHop + i;
                        if (j < 0) j = -j;
                        if (j >= N) j = 2 * (N - 1) - j;
                        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
                        raw[g * kRawGroupPitch + i] = __ldg(x + j);
                    }
                }
                named_bar_sync(2, 512);
            }
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 27);
            for (int ks = kpar; ks < kKSteps; ks += 2) {
                const uint32_t it = tl * kKSteps + ks;
                const int sa = it % kAStages;
                float hi[8], lo[8];
                if (!(p.dbg & 2)) {
                    float v[8];
                    const int n0 = ks * 8 + 1;                           // n = n0 + jj in 1..200
                    const float* rev_row = (kNfft - n0 - 7 >= 2 * kHop) ? row_d2 : row_d1;   // m = 400 - n: 8-aligned windows never straddle 320
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int n = n0 + jj[j];
                        const float fwd = (n >= kHop ? row_d1 : row_d0)[n];
                        const float rev = rev_row[kNfft - n];
                        v[j] = odd ? fwd - rev : fwd + rev;              // n = 200: e = 2 x[200] (basis row halved), o = 0
                    }
                    // undo the rotation: w[c] = v[(c - rot) & 7]
                    float t[8];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) t[cidx] = (rot & 1) ? v[(cidx + 7) & 7] : v[cidx];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) v[cidx] = (rot & 2) ? t[(cidx + 6) & 7] : t[cidx];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) t[cidx] = (rot & 4) ? v[(cidx + 4) & 7] : v[cidx];
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) {
                        hi[cidx] = rna_tf32(t[cidx]);
                        lo[cidx] = rna_tf32(t[cidx] - hi[cidx]);
                    }
                }
                if (warp == 12 && lane == 0 && ks == 0) trace(p.dbg, 1, tl, 28);
                // only the stores need the TMEM stage: everything above overlaps the MMAs in flight
                mbar_wait(&a_empty[sa], ((it / kAStages) & 1) ^ 1);
                tc_fence_after();
                if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 1 + ks);
                if (!(p.dbg & 2)) {
                    tmem_st8(tlane + sa * 32, hi);
                    tmem_st8(tlane + sa * 32 + 8, lo);
                    tmem_st_wait();
                }
                if (warp == 12 && lane == 0 && ks == 0) trace(p.dbg, 1, tl, 29);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[sa]);
                if (warp == 12 && lane == 0 && ks == 0) trace(p.dbg, 1, tl, 30);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(raw_empty);      // staged waveform no longer needed
            if (warp == 12 && lane == 0) trace(p.dbg, 1, tl, 26);
        }
    } else if (warp >= 4 && warp < 12) {
        // ====================== epilogue: power -> mel, straight out of TMEM (256 threads) ====
        // Two warps share each TMEM lane quarter and split the 201 bins at kSplit. Exactly two
        // filters (ms, ms + 1, ms = first filter fed by bin kSplit) receive power from both sides.
        // The mel POWER is stored; log10 / floor / scaling happen in logmel_finalize_kernel (max is
        // monotone under log10). The loop body is kept small on purpose: the first version unrolled
        // to 3200 instructions and spent its time in instruction fetch.
        const int half = warp >= 8 ? 1 : 0;
        const int wq = warp & 3;                   // TMEM lane quarter == warp % 4
        const int row = wq * 32 + lane;
        constexpr int ms = LA_MEL_MS;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
            const ClipDesc c = p.clips[p.tile_clip[tile]];
            const int f0 = (tile - c.tile0) * kTileM;
            mbar_wait(tmem_full, tl & 1);
            tc_fence_after();
            if (warp == 4 && lane == 0) trace(p.dbg, 2, tl, 0);
            const int f = f0 + row;
            const bool valid = f < c.n_frames;
            const int64_t ostride = c.out_stride;
            float* outp = p.out + c.out_off + f;
            float* optr = outp + (half ? ms : 0) * ostride;   // where the filter held in a0 will be stored
            float mx = 0.f;
            float a0 = 0.f, a1 = 0.f;
            auto flush = [&]() {                    // the filter held in a0 is complete: store its power
                if (valid) { *optr = a0; mx = fmaxf(mx, a0); }   // rows past the clip's last frame hold garbage
                optr += ostride;
                a0 = a1; a1 = 0.f;
            };
            const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
            uint32_t re[16], im[16];
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
                const float xr = __uint_as_float(re[(K) & 15]), xi = __uint_as_float(im[(K) & 15]); \
                const float pw = xr * xr + xi * xi;                                               \
                if ((NF) >= 1 && (K) != 0 && (K) != kSplit) flush();                              \
                if ((NF) >= 2 && (K) != 0 && (K) != kSplit) flush();                              \
                a0 = fmaf(W0, pw, a0);                                                            \
                a1 = fmaf(W1, pw, a1);                                                            \
            }
#define LA_CHUNK(C, LAST)                                                                         \
            {                                                                                     \
                tmem_ld16(lane_base + 16 * (C), re);                                              \
                tmem_ld16(lane_base + kNpad + 16 * (C), im);                                      \
                tmem_ld_wait();                                                                   \
                if (LAST) release_tmem();                                                         \
                LA_MEL_CHUNK_##C(LA_BIN)                                                          \
            }
            if (p.dbg & 1) {
                release_tmem();
                continue;
            }
            if (!half) {
                LA_CHUNK(0, false) LA_CHUNK(1, false) LA_CHUNK(2, false)
                LA_CHUNK(3, false) LA_CHUNK(4, false) LA_CHUNK(5, true)
#pragma unroll
                for (int i = 0; i < LA_MEL_TAIL0; ++i) flush();
            } else {
                LA_CHUNK(6, false) LA_CHUNK(7, false) LA_CHUNK(8, false) LA_CHUNK(9, false)
                LA_CHUNK(10, false) LA_CHUNK(11, false) LA_CHUNK(12, true)
#pragma unroll
                for (int i = 0; i < LA_MEL_TAIL1; ++i) flush();
            }
#undef LA_CHUNK
#undef LA_BIN
            if ((warp == 4 || warp == 8) && lane == 0) trace(p.dbg, 2, tl, warp == 4 ? 3 : 6);
            // Filters ms and ms + 1 are fed from both halves: the high half stores its partial
            // sums like any other filter, the low half adds its own after the barrier.
            named_bar_sync(3, 256);                     // both halves, one barrier instruction
            if (warp == 4 && lane == 0) trace(p.dbg, 2, tl, 4);
            if (!half && valid) {
                float* o0 = outp + ms * ostride;
                const float v0 = *o0 + a0, v1 = o0[ostride] + a1;
                *o0 = v0;
                o0[ostride] = v1;
                mx = fmaxf(mx, fmaxf(v0, v1));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) atomicMax(p.group_max + c.group, __float_as_int(mx));   // mx >= 0: int order == float order
            if (warp == 4 && lane == 0) trace(p.dbg, 2, tl, 1);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// mel power -> log10(clamp 1e-10) -> max(x, gmax - 8) -> (x + 4) / 4 over the valid frames of every clip
__global__ void logmel_finalize_kernel(const LogmelParams p) {
    const ClipDesc c = p.clips[blockIdx.y];
    const int f = blockIdx.x * blockDim.x + threadIdx.x;       // one frame column per thread, coalesced along f
    if (f >= c.n_frames) return;
    const float gmax = log10f(fmaxf(__int_as_float(p.group_max[c.group]), 1e-10f));
    const float floor_v = gmax - 8.0f;
    float* q = p.out + c.out_off + f;
#pragma unroll 8
    for (int m = 0; m < kMels; ++m, q += c.out_stride) {
        const float x = log10f(fmaxf(*q, 1e-10f));
        *q = (fmaxf(x, floor_v) + 4.0f) / 4.0f;
    }
}

__global__ void fill_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- host: constant tables ---------------------------------------------------------------
static void tf32_split(double v, float* hi, float* lo) {
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u = (u + 0x1000u) & 0xffffe000u;          // round to nearest TF32 (10 explicit mantissa bits)
    float h;
    memcpy(&h, &u, 4);
    *hi = h;
    float l = (float)(v - (double)h);
    memcpy(&u, &l, 4);
    u = (u + 0x1000u) & 0xffffe000u;          // pre-round lo as well: operand truncation becomes a no-op
    memcpy(&l, &u, 4);
    *lo = l;
}

struct LogmelTables {
    std::mutex mu;
    float* d_basis[64] = {nullptr};
};
static LogmelTables g_tab;

static cudaError_t ensure_tables(int device, const float** basis_out) {
    std::lock_guard<std::mutex> lock(g_tab.mu);
    if (!g_tab.d_basis[device]) {
        // basis blocks: per k-step j: [C_hi | C_lo] (even half-step) then [S_hi | S_lo] (odd), each
        // matrix in the canonical K-major layout [ki 2][ni 26][8 rows (bin)][4 floats (sample)]
        std::vector<float> h((size_t)kKSteps * 4 * (kBBytes / 4), 0.f);
        const double PI = 3.14159265358979323846;
        for (int j = 0; j < kKSteps; ++j)
            for (int kk = 0; kk < 8; ++kk) {
                const int n = j * 8 + kk + 1;                               // 1..200
                const double w = 0.5 - 0.5 * std::cos(2.0 * PI * n / kNfft);
                for (int b = 0; b < kNpad; ++b) {
                    double cv = 0.0, sv = 0.0;
                    if (b < kBins) {
                        const int ph = (int)(((long long)b * n) % kNfft);     // exact phase reduction
                        // n = 200 folds onto itself: the kernel forms e[200] = 2 x[200], so halve its row
                        cv = (n == kNfft / 2 ? 0.5 : 1.0) * w * std::cos(2.0 * PI * ph / kNfft);
                        sv = (n == kNfft / 2) ? 0.0 : -w * std::sin(2.0 * PI * ph / kNfft);
                    }
                    const size_t inner = (size_t)(kk >> 2) * (kBLbo / 4) + (size_t)(b >> 3) * 32 + (b & 7) * 4 + (kk & 3);
                    const size_t base = (size_t)j * 4 * (kBBytes / 4);
                    tf32_split(cv, &h[base + inner], &h[base + (kBBytes / 4) + inner]);
                    tf32_split(sv, &h[base + 2 * (kBBytes / 4) + inner], &h[base + 3 * (kBBytes / 4) + inner]);
                }
            }
        float* d = nullptr;
        cudaError_t e = cudaMalloc(&d, h.size() * 4);
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(d); return e; }
        g_tab.d_basis[device] = d;
    }
    *basis_out = g_tab.d_basis[device];
    return cudaSuccess;
}

size_t logmel_smem_bytes() { return (size_t)kBStages * kBStageBytes + kRawBytes + 256; }

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
    const float* basis = nullptr;
    e = ensure_tables(device, &basis);
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "tables: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    std::vector<ClipDesc> clips = clips_in;
    std::vector<int32_t> tile_clip;
    int max_frames = 0;
    for (size_t ci = 0; ci < clips.size(); ++ci) {
        clips[ci].tile0 = (int32_t)tile_clip.size();
        const int nt = (clips[ci].n_frames + kTileM - 1) / kTileM;
        for (int t = 0; t < nt; ++t) tile_clip.push_back((int32_t)ci);
        max_frames = std::max(max_frames, clips[ci].n_frames);
    }
    const int n_tiles = (int)tile_clip.size();
    if (n_tiles == 0) return LA_OK;
    const size_t o_clips = 0;
    const size_t o_tiles = lm_align(o_clips + clips.size() * sizeof(ClipDesc), 256);
    const size_t o_max = lm_align(o_tiles + tile_clip.size() * 4, 256);
    const size_t need = o_max + lm_align((size_t)n_groups * 4, 256);
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
    { const char* d = getenv("LA_LOGMEL_DBG"); p.dbg = d ? atoi(d) : 0; }
    fill_int_kernel<<<(n_groups + 255) / 256, 256, 0, stream>>>(p.group_max, n_groups, 0);   // 0.0f: powers are >= 0
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const size_t smem = logmel_smem_bytes();
    e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "smem attr: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    const int ctas = la::g_logmel_ctas > 0 ? std::min(la::g_logmel_ctas, sms) : sms;
    logmel_kernel<<<std::min(n_tiles, ctas), kLogmelThreads, smem, stream>>>(p);
    for (size_t c0 = 0; c0 < clips.size(); c0 += 65535) {            // gridDim.y limit
        LogmelParams q = p;
        q.clips = p.clips + c0;
        const unsigned ny = (unsigned)std::min<size_t>(65535, clips.size() - c0);
        logmel_finalize_kernel<<<dim3((max_frames + 127) / 128, ny), 128, 0, stream>>>(q);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_lm_err, sizeof g_lm_err, "launch: %s", cudaGetErrorString(e)); LM_FAIL(LA_ERR_CUDA); }
    return LA_OK;
}

extern "C" {

// perf triage only (not part of the public header): copies the LA_LOGMEL_DBG&8 timestamps out
int la_debug_logmel_trace(unsigned long long* h_out) {
    return cudaMemcpyFromSymbol(h_out, la::g_trace, sizeof(unsigned long long) * 4 * 8 * 32) == cudaSuccess ? 0 : -2;
}

size_t la_logmel_workspace_bytes(int n_clips, int64_t total_samples) {
    if (n_clips < 0 || total_samples < 0) return 0;
    const size_t tiles = (size_t)(total_samples / la::kHop) / la::kTileM + (size_t)n_clips + 1;
    return lm_align((size_t)n_clips * sizeof(la::ClipDesc), 256) + lm_align(tiles * 4, 256) +
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
