// la_viterbi.cu -- K3: Viterbi forced-alignment DP + 2-bit packed backpointers + backtrace.
//
// Replaces run_viterbi_core (utils/alignment.py:73-119), the dp/bt allocation (:144-152), the
// strict end-state pick and backtrace (:157-176) and the first/last on/offset scan (:182-185).
//
// Bit-exactness contract: state is IEEE fp64, one DADD per cell on an exactly
// promoted fp32 emission; the comparisons are the reference's own (`>` strict for stay vs
// previous state, `>=` for the label skip) evaluated in its branch order; EVERY state of EVERY
// frame is computed (no pruning); unreachable cells start at the finite floor -1e7.
//
// Mapping: the 2L+1 states are grouped in L+1 "pairs" -- pair i = (blank state 2i, label state
// 2i+1). A pair only needs ONE value from its left neighbour (the previous label's score of the previous frame).
// Two kernels share that mapping, the backpointer format and the backtrace:
// * la_viterbi_wave.cuh -- the WAVEFRONT kernel, utterances of up to 639 pairs (every BASELINE config): lane i works
//   i frames behind lane 0, so the neighbour's value is a step old when it is needed and the shuffle leaves the
//   dependent chain. One warp up to 63 pairs, a pipeline of up to ten warps beyond. See that file.
// * viterbi_kernel below -- the ROW-SYNCHRONOUS kernel, 640 .. 8192 pairs: all lanes work on the same frame, a thread
//   owns K = 2/4/8 consecutive pairs and needs one 64-bit warp shuffle per frame; one CTA per utterance with exactly
//   ceil(pairs / 32K) warps. The dependency runs left to right only, so warps are a PIPELINE, not a lock-step team:
//   warp w hands the score of its last pair to warp w+1 through a shared-memory ring of 2 x chunk slots (a 64-bit
//   store; the consumer polls an all-ones NaN sentinel and re-arms the slot). No block barrier in the frame loop and
//   none per chunk either: the emission stages are recycled through full/empty mbarriers, so the warps never
//   re-align and the pipeline never drains. Emission rows ([T][row_floats], col 0 = blank) are streamed in chunks of
//   up to 32 frames into a 2- to 4-deep shared-memory ring by 1-D TMA bulk copies issued by the last warp.
// Backpointers are 2-bit step codes (k - bt), one nibble per pair per frame, 8 frames per 32-bit word, written
// coalesced. The backtrace (backtrace_walk) is a single warp searching shared-memory tiles of those words for the next
// transition. Lane 0 emits first / last+1 at every label-state run boundary (the path is monotone, so each label's
// occupancy is one run).
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "la_common.cuh"

namespace la {

constexpr int kVitChunkMax = 32;   // frames per TMA chunk (fewer when rows are very wide)
constexpr int kVitStagesMax = 4;
__device__ unsigned long long g_vit_trace[4];   // LA_VIT_TRACE=1: CTA 0's clock64 at DP start / DP end / backtrace end, and T
constexpr unsigned long long kNotReady = ~0ull;   // all-ones NaN: neither arithmetic nor an f32->f64 promotion can produce it

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}
// Hand-off ring (shared-memory byte addresses). take: EVERY lane of the warp polls the same slot (one
// broadcast LDS, the loop exit is warp-uniform, no divergence); lane 0 then re-arms it. put: a predicated
// store, no branch. No "memory" clobber: the ring is ordered by `volatile` among these asm statements
// only, so the compiler stays free to hoist the emission loads above them.
__device__ __forceinline__ double ring_take(uint32_t addr, uint32_t rearm) {
    unsigned long long v;
    do {
        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    } while (v == kNotReady);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.volatile.shared.u64 [%0], %1;\n\t}"
                 ::"r"(addr), "l"(kNotReady), "r"(rearm));
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void ring_put(uint32_t addr, double v, uint32_t pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.volatile.shared.u64 [%0], %1;\n\t}"
                 ::"r"(addr), "l"(__double_as_longlong(v)), "r"(pred));
}

// ---- backtrace: one warp; a tile of 32 word rows (256 frames) x 32 columns of packed codes in shared memory --------
// The walker only ever moves to smaller t and smaller k. A state is usually occupied for many frames, so the walk is
// a search: "the latest frame <= t at which state k's code is non-zero". Lane r reads the word of tile row r in the
// walker's column, masks it (state type, frames behind the walker, frame 0), the warp votes, and the walker jumps
// straight to the next transition anywhere in the tile's 256 frames: one iteration per transition or per tile. A lone
// warp issues a dependent instruction every ~5.5 cycles, so the iteration is written as the shortest chain that does
// the job: every lane prepares its own candidate (frame << 2 | step) and ONE warp-wide max (REDUX) picks the latest;
// the step is read off the position of the highest set bit (blank states:
// bit 0 of a nibble = step 1; label states: bit 1 = step 1, bit 2 = step 2).
// Tiles are fetched with 16-byte asynchronous copies; the NEXT tile (same columns, the 32 word rows before) is requested
// as soon as the current one is entered, which hides the L2 round trip in the time direction. An utterance of up to 32
// columns never leaves its tile sideways. Layout: pair i lives in column i + SH, nibble t%8 of word row t/8 is frame t.
// `tiles`: 2 x kBtTileWords words of shared memory (16-byte aligned) the DP no longer needs.
// Returns the number of label states visited (== L iff every label is on the path).
constexpr int kBtPitch = 36;                              // words per tile row: 32 columns + 4 (16-byte multiple, 4-way banks)
constexpr int kBtTileWords = 32 * kBtPitch;
constexpr size_t kBtSmemBytes = 2 * kBtTileWords * 4;

template <int SH>
__device__ __forceinline__ int backtrace_walk(const uint32_t* __restrict__ bp, int pairs_pad, int T, int k, int lane,
                                              int32_t* __restrict__ first, int32_t* __restrict__ lastp, uint32_t* tiles) {
    int visited = 0;
    if ((k & 1) && lane == 0) lastp[k >> 1] = T;
    const uint32_t tiles_u32 = smem_u32(tiles);
    // tile `buf` <- word rows b0-31 .. b0 (row b0 first), columns p0 .. p0+31; one commit group
    auto fetch = [&](int buf, int b0, int p0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int id = j * 32 + lane, rr = id >> 3, c4 = (id & 7) * 4;
            const int row = b0 - rr;
            const bool ok = row >= 0 && p0 + c4 < pairs_pad;
            const uint32_t dst = tiles_u32 + (uint32_t)(buf * kBtTileWords + rr * kBtPitch + c4) * 4u;
            const uint32_t* src = bp + (int64_t)(ok ? row : 0) * pairs_pad + (ok ? p0 + c4 : 0);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                         ::"r"(dst), "l"(src), "r"((uint32_t)ok) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    asm volatile("" : "+l"(first), "+l"(lastp));          // keep both pointers in registers (else: an LDC per iteration)
    int t = T - 1;
    int cur = 0, b0 = t >> 3, p0 = max(0, (((k >> 1) + SH) | 3) - 31);
    __syncwarp();
    fetch(0, b0, p0);
    fetch(1, b0 - 32, p0);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    // per lane, per tile: first frame of this lane's word, its address inside the tile, and the nibbles that may hold a
    // code at all (row 0: frame 0 carries none; rows < 0 were never fetched)
    int base = (b0 - lane) << 3;
    uint32_t live = base > 0 ? 0xffffffffu : (base == 0 ? 0xfffffff0u : 0u);
    const uint32_t* trow = tiles + lane * kBtPitch - p0;
    while (t >= 1) {
        const int col = (k >> 1) + SH;
        if ((unsigned)(b0 - (t >> 3)) >= 32u || (unsigned)(col - p0) >= 32u) {
            __syncwarp();                                 // every lane has read the tile that is about to be replaced
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if ((unsigned)(b0 - 32 - (t >> 3)) < 32u && (unsigned)(col - p0) < 32u) {   // the prefetched tile
                b0 -= 32;
                fetch(cur, b0 - 32, p0);
                cur ^= 1;
            } else {                                      // left the tile sideways (or jumped past a whole tile)
                b0 = t >> 3;
                p0 = max(0, (col | 3) - 31);
                fetch(cur, b0, p0);
                fetch(cur ^ 1, b0 - 32, p0);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
            __syncwarp();
            base = (b0 - lane) << 3;
            live = base > 0 ? 0xffffffffu : (base == 0 ? 0xfffffff0u : 0u);
            trow = tiles + cur * kBtTileWords + lane * kBtPitch - p0;
        }
        // codes of state k in this lane's word, frames <= t only (the rest is behind the walker)
        const uint32_t kind = ((k & 1) ? 0x66666666u : 0x11111111u) & live;
        const int hi = t - base;                          // nibbles 0 .. hi are in range
        const uint32_t upto = hi >= 7 ? 0xffffffffu : (hi < 0 ? 0u : (0xffffffffu >> (28 - 4 * hi)));
        const uint32_t m = trow[col] & kind & upto;
        const int pbit = 31 - __clz(m);                   // highest code bit of this lane's word (if any)
        const uint32_t cand = m ? ((uint32_t)((base + (pbit >> 2)) << 2) | (uint32_t)((pbit & 3) ? (pbit & 3) : 1)) : 0u;
        const uint32_t win = __reduce_max_sync(0xffffffffu, cand);   // one REDUX: the latest frame any lane found
        if (win == 0u) {                                  // state k stays down to the bottom of the tile
            t = ((b0 - 31) << 3) - 1;
            continue;
        }
        t = (int)(win >> 2);                              // latest frame <= t with a non-zero code
        if (k & 1) {                                       // label state k occupied frames t..: onset
            if (lane == 0) first[k >> 1] = t;
            ++visited;
        }
        k -= (int)(win & 3u);
        if ((k & 1) && lane == 0) lastp[k >> 1] = t;       // new label state ends at frame t-1
        --t;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // nothing may still be writing shared memory at exit
    if (k & 1) {
        if (lane == 0) first[k >> 1] = 0;
        ++visited;
    }
    return visited;
}

// K = pairs per thread, DUMP = parity instrumentation (full fp64 table to global memory; compiled out of
// the production kernels)
template <int K, bool DUMP, int MAXT>
__global__ void __launch_bounds__(MAXT) viterbi_kernel(const VitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw_launch = blockDim.x >> 5;

    // ---- shared memory carve-up: stages, full/empty mbarriers, 2 finals, hand-off ring ----
    const int chunk = p.chunk, stages = p.stages;
    const int ring = p.ring;                              // slots per warp boundary (power of two >= stages * chunk)
    const int stage_bytes = chunk * p.row_floats_max * 4;
    float* stage0 = reinterpret_cast<float*>(smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);     // [kVitStagesMax]
    uint64_t* empty = full + kVitStagesMax;                                         // [kVitStagesMax]
    double* fin = reinterpret_cast<double*>(empty + kVitStagesMax);                 // [2]
    unsigned long long* xchg = reinterpret_cast<unsigned long long*>(fin + 2);      // [nw_launch][ring]

    const int utt = p.order[blockIdx.x];
    const int T = p.m.t_off[utt + 1] - p.m.t_off[utt];
    const int l0 = p.m.l_off[utt];
    const int L = p.m.l_off[utt + 1] - l0;
    if (L <= 0 || T <= 0) {                               // uniform per CTA
        if (tid == 0) {
            p.status[utt] = (L <= 0) ? 1 : 2;
            p.score[utt] = 0.0;
        }
        return;
    }
    const int wrow = p.m.e_row[utt];
    const float* E = p.E + p.m.e_off[utt];
    const int pairs_pad = p.m.bp_pairs[utt];
    uint32_t* bp = p.bp + p.m.bp_off[utt];
    // the launch is sized for the widest utterance of the bucket; warps this one does not need
    // leave before the first barrier (exited warps do not take part in __syncthreads)
    const int nwarps = pairs_pad / (32 * K);
    if (warp >= nwarps) return;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
        mbar_fence_init();
    }
    for (int i = tid; i < ring * nw_launch; i += 32 * nwarps) xchg[i] = kNotReady;
    __syncthreads();

    const int nchunks = (T + chunk - 1) / chunk;
    auto issue = [&](int c, int s) {                      // chunk c into stage s (= c % stages; the callers track it)
        const int rows = min(chunk, T - c * chunk);
        const uint32_t bytes = (uint32_t)rows * wrow * 4;
        fence_proxy_async();
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(reinterpret_cast<unsigned char*>(stage0) + s * stage_bytes, E + (int64_t)c * chunk * wrow, bytes, &full[s]);
    };
    // The LAST warp refills the stages: it is the pipeline's laggard, so by the time it starts chunk c every
    // other warp has (all but certainly) left chunk c-1 and the wait on empty[] costs nothing, while the
    // leading warp -- up to nwarps frames ahead -- still has stages-1 chunks of emissions staged.
    const bool loader = (warp == nwarps - 1) && lane == 0;
    if (loader)
        for (int c = 0; c < min(stages, nchunks); ++c) issue(c, c);

    // ---- per-pair constants ------------------------------------------------------------
    const int pair0 = tid * K;
    int ecol[K];
    bool skip_ok[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = pair0 + j;
        ecol[j] = (i < L) ? 1 + i : 0;                    // dummy / last pair: any valid column
        skip_ok[j] = (i >= 1 && i < L) ? (p.m.labels[l0 + i] != p.m.labels[l0 + i - 1]) : false;
    }
    double pb[K], pl[K];
    uint32_t acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { pb[j] = kFloor; pl[j] = kFloor; acc[j] = 0u; }
    const bool has_left = warp > 0;                        // warp-uniform
    const bool lane0 = lane == 0;
    uint32_t put_pred = (lane == 31 && warp + 1 < nwarps) ? 1u : 0u;
    uint32_t rearm = lane0 ? 1u : 0u;
    uint32_t put_base = smem_u32(xchg) + (uint32_t)(warp * ring) * 8u;          // my slots, read by warp + 1
    uint32_t take_base = smem_u32(xchg) + (uint32_t)(max(warp, 1) - 1) * ring * 8u;   // the left neighbour's slots
    const uint32_t ring_mask8 = (uint32_t)(ring - 1) * 8u;
    // opaque to the optimiser: otherwise ptxas re-derives these from S2R/cvta inside the frame loop
    asm volatile("" : "+r"(put_pred), "+r"(rearm), "+r"(put_base), "+r"(take_base));

    // one frame, any K, every special case checked (frame 0 preset, tail word, DUMP)
    auto step = [&](int t, double eb, const double (&el)[K]) {
        if (t == 0) {
            // row 0 preset (utils/alignment.py:151-152)
            if (tid == 0) { pb[0] = eb; pl[0] = el[0]; }
            ring_put(put_base, pl[K - 1], put_pred);
        } else {
            double ql = shfl_up_f64(pl[K - 1], 1);
            if (has_left) {
                const double v = ring_take(take_base + (((uint32_t)(t - 1) * 8u) & ring_mask8), rearm);
                ql = lane0 ? v : ql;
            } else {
                ql = lane0 ? -INFINITY : ql;               // pair 0: no left neighbour
            }
            const int sh = (t & 7) * 4;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double b = pb[j], l = pl[j];
                const double q = ql;
                ql = l;                                    // left neighbour of pair j+1 (old value)
                // blank state 2i (utils/alignment.py:78-82, 92-101)
                const bool b_stay = b > q;
                pb[j] = (b_stay ? b : q) + eb;
                // label state 2i+1 (:84-90, 103-117)
                const bool skip = (q >= b) && (q >= l) && skip_ok[j];
                const bool l_stay = l > b;
                pl[j] = (skip ? q : (l_stay ? l : b)) + el[j];
                const uint32_t nib = (b_stay ? 0u : 1u) | (skip ? 4u : (l_stay ? 0u : 2u));
                acc[j] |= nib << sh;
            }
            ring_put(put_base + (((uint32_t)t * 8u) & ring_mask8), pl[K - 1], put_pred);
        }
        if (DUMP) {
            double* drow = p.dp_dump + (int64_t)t * (2 * L + 1);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (pair0 + j <= L) drow[2 * (pair0 + j)] = pb[j];
                if (pair0 + j < L) drow[2 * (pair0 + j) + 1] = pl[j];
            }
        }
        if ((t & 7) == 7 || t == T - 1) {
            uint32_t* w = bp + (int64_t)(t >> 3) * pairs_pad + pair0;
            if (K == 4) {
                *reinterpret_cast<uint4*>(w) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            } else if (K == 2) {
                *reinterpret_cast<uint2*>(w) = make_uint2(acc[0], acc[1]);
            } else {
#pragma unroll
                for (int j = 0; j < K; ++j) w[j] = acc[j];
            }
#pragma unroll
            for (int j = 0; j < K; ++j) acc[j] = 0u;
        }
    };
    // eight frames t .. t+7 (t a multiple of 8, t >= 8, all inside one chunk), K <= 2 pairs per lane: the
    // emissions of the whole block are in registers up front, shifts and ring offsets are immediates, and the
    // only per-frame work left is the dependent chain itself (two independent ones per lane when K = 2, which is
    // what lets a lone warp overlap their latencies)
    auto block8 = [&](int t, const float* r0, auto HL) {
        constexpr bool kHasLeft = decltype(HL)::value;
        float ebf[8], elf[8][K];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ebf[i] = r0[i * wrow];
#pragma unroll
            for (int j = 0; j < K; ++j) elf[i][j] = r0[i * wrow + ecol[j]];
        }
        const uint32_t put8 = put_base + (((uint32_t)t * 8u) & ring_mask8);            // slots t .. t+7 are contiguous
        const uint32_t take0 = take_base + (((uint32_t)(t - 1) * 8u) & ring_mask8);    // slot t-1
        const uint32_t take8 = take_base + (((uint32_t)t * 8u) & ring_mask8);          // slots t .. t+6 feed frames t+1 .. t+7
        uint32_t a[K];
        double b[K], l[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { a[j] = 0u; b[j] = pb[j]; l[j] = pl[j]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double q = shfl_up_f64(l[K - 1], 1);
            if (kHasLeft) {
                const double v = ring_take(i == 0 ? take0 : take8 + 8u * (uint32_t)(i - 1), rearm);
                q = lane0 ? v : q;
            } else {
                q = lane0 ? -INFINITY : q;
            }
            const double eb = (double)ebf[i];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double bj = b[j], lj = l[j];
                const bool b_stay = bj > q;
                const bool skip = (q >= bj) && (q >= lj) && skip_ok[j];
                const bool l_stay = lj > bj;
                b[j] = (b_stay ? bj : q) + eb;
                l[j] = (skip ? q : (l_stay ? lj : bj)) + (double)elf[i][j];
                a[j] |= ((b_stay ? 0u : 1u) | (skip ? 4u : (l_stay ? 0u : 2u))) << (4 * i);
                q = lj;                                    // left neighbour of pair j+1 (old value)
            }
            ring_put(put8 + 8u * (uint32_t)i, l[K - 1], put_pred);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) { pb[j] = b[j]; pl[j] = l[j]; }
        uint32_t* w = bp + (int64_t)(t >> 3) * pairs_pad + pair0;
        if (K == 2) *reinterpret_cast<uint2*>(w) = make_uint2(a[0], a[1]);
        else w[0] = a[0];
    };

    const long long c_fwd0 = clock64();
    int s = 0, ph = 0, sp = 0, php = 0;                    // stage / phase parity of chunk c and of chunk c-1 (no div/mod per chunk)
    for (int c = 0; c < nchunks; ++c) {
        if (loader && c >= 1 && c - 1 + stages < nchunks) {
            mbar_wait(&empty[sp], php);                    // every warp has left chunk c-1
            issue(c - 1 + stages, sp);
        }
        mbar_wait(&full[s], ph);
        const float* rows = stage0 + s * (stage_bytes / 4);
        const int t0 = c * chunk;
        const int nt = min(chunk, T - t0);
        int tt = 0;
        while (tt < nt) {
            const int t = t0 + tt;
            if (K <= 2 && !DUMP && (t & 7) == 0 && t >= 8 && tt + 8 <= nt) {
                if (has_left) block8(t, rows + tt * wrow, std::true_type{});
                else block8(t, rows + tt * wrow, std::false_type{});
                tt += 8;
                continue;
            }
            const float* er = rows + tt * wrow;
            double el[K];
#pragma unroll
            for (int j = 0; j < K; ++j) el[j] = (double)er[ecol[j]];
            step(t, (double)er[0], el);
            ++tt;
        }
        __syncwarp();
        if (lane0) mbar_arrive(&empty[s]);                 // this warp no longer reads stage s
        sp = s; php = ph;
        if (++s == stages) { s = 0; ph ^= 1; }
    }

    // ---- end-state pick (utils/alignment.py:157): S-1 iff dp[T-1][S-1] > dp[T-1][S-2] --------
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (pair0 + j == L) fin[0] = pb[j];
        if (pair0 + j == L - 1) fin[1] = pl[j];
    }
    __syncthreads();                                       // all warps done; also orders their bp stores before the walker's loads
    if (warp != 0) return;

    const long long c_fwd1 = clock64();
    int k = (fin[0] > fin[1]) ? 2 * L : 2 * L - 1;
    const double best = (fin[0] > fin[1]) ? fin[0] : fin[1];

    const int visited = backtrace_walk<0>(bp, pairs_pad, T, k, lane, p.first + l0, p.last_plus1 + l0,
                                                    reinterpret_cast<uint32_t*>(smem));
    if (lane == 0) {
        p.status[utt] = (visited == L) ? 0 : 2;          // a missing label state -> ValueError upstream
        p.score[utt] = best;
        if (p.trace && blockIdx.x == 0) {
            g_vit_trace[0] = (unsigned long long)(c_fwd0);
            g_vit_trace[1] = (unsigned long long)(c_fwd1);
            g_vit_trace[2] = (unsigned long long)clock64();
            g_vit_trace[3] = (unsigned long long)T;
        }
    }
}

#include "la_viterbi_wave.cuh"

// frames per chunk: a power of two (8-frame blocks and the hand-off ring index with masks), at most 32. The
// per-chunk bookkeeping (barrier waits, refill) is worth ~700 cycles, so wide rows get the biggest chunk of which
// TWO stages still fit in 160 KB (the refill distance is then one chunk = 32 frames = ~6 us, far more than the
// copy needs), narrow rows keep 32-frame chunks with three or four stages.
int viterbi_chunk_frames(int row_floats_max) {
    const int fit = (80 * 1024) / (row_floats_max * 4);
    int c = kVitChunkMax;
    while (c > 1 && c > fit) c >>= 1;
    return c;
}
// the leading warp runs up to `warps` frames ahead of the last one, and a stage is only refilled once every warp
// has left it: (stages - 1) chunks must cover that skew
static int viterbi_stages(int warps, int chunk, int row_floats_max) {
    const int stage_bytes = chunk * row_floats_max * 4;
    const int fit = std::max(2, std::min(kVitStagesMax, (160 * 1024) / std::max(stage_bytes, 1)));
    const int want = warps > 4 ? 4 : 3;
    int st = std::min(want, fit);
    while (st < kVitStagesMax && (st - 1) * chunk < warps + 8 && (st + 1) * stage_bytes <= 200 * 1024) ++st;
    return st;
}
static int viterbi_ring(int chunk, int stages) {
    int r = 16;
    while (r < stages * chunk) r <<= 1;
    return r;
}

size_t viterbi_smem_bytes(int row_floats_max, int chunk, int stages, int ring, int warps) {
    return (size_t)stages * chunk * row_floats_max * 4 + 2 * kVitStagesMax * 8 + 16 + (size_t)ring * warps * 8;
}

template <int K, int MAXT>
static cudaError_t launch_one(const VitParams& p, int threads, int grid, size_t smem, cudaStream_t stream) {
    // the opt-in is a per-function attribute of the loaded module: set once per (function, device), not per launch
    static bool attr_done[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int d = p.dp_dump ? 1 : 0;
    if (dev >= 0 && dev < 64 && !attr_done[d][dev]) {
        cudaError_t e = p.dp_dump
            ? cudaFuncSetAttribute(viterbi_kernel<K, true, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)
            : cudaFuncSetAttribute(viterbi_kernel<K, false, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return e;
        attr_done[d][dev] = true;
    }
    if (p.dp_dump) viterbi_kernel<K, true, MAXT><<<grid, threads, smem, stream>>>(p);
    else viterbi_kernel<K, false, MAXT><<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

// K pairs per lane; one CTA of `warps` warps per utterance. CTAs of up to 16 warps are compiled with a 512-thread
// bound (128 registers: the unrolled 8-frame block keeps 8 x (1 + K) emissions live), the 32-warp shapes with 1024.
cudaError_t launch_viterbi(const VitParams& p_in, int K, int warps, cudaStream_t stream) {
    if (p_in.n_order <= 0) return cudaSuccess;
    VitParams p = p_in;
    static const int trace = [] { const char* e = getenv("LA_VIT_TRACE"); return e ? atoi(e) : 0; }();
    p.trace = trace;
    p.stages = viterbi_stages(warps, p.chunk, p.row_floats_max);
    p.ring = viterbi_ring(p.chunk, p.stages);
    const size_t smem = viterbi_smem_bytes(p.row_floats_max, p.chunk, p.stages, p.ring, warps);
    const int threads = 32 * warps;
    if (warps <= 16 && K == 2) return launch_one<2, 512>(p, threads, p.n_order, smem, stream);
    switch (K) {
        case 2: return launch_one<2, 1024>(p, threads, p.n_order, smem, stream);
        case 4: return launch_one<4, 1024>(p, threads, p.n_order, smem, stream);
        default: return launch_one<8, 1024>(p, threads, p.n_order, smem, stream);
    }
}

}  // namespace la

// perf triage only (not part of the public header)
extern "C" int la_debug_viterbi_trace(unsigned long long* h_out) {
    return cudaMemcpyFromSymbol(h_out, la::g_vit_trace, sizeof(unsigned long long) * 4) == cudaSuccess ? 0 : -2;
}
