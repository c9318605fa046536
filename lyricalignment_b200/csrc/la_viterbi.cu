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
// 2i+1). A pair only needs ONE value from its left neighbour (the previous label's score), so a
// thread that owns K consecutive pairs needs one 64-bit warp shuffle per frame. One CTA per
// utterance with exactly ceil(pairs / 32K) warps. The dependency runs left to right only, so warps
// are a PIPELINE, not a lock-step team: warp w hands the score of its last pair to warp w+1
// through a shared-memory ring of 2 x chunk slots (a 64-bit store; the consumer's lane 0 spins on
// an all-ones NaN sentinel and re-arms the slot). There is NO block barrier in the frame loop --
// round 1's per-frame __syncthreads cost 160-250 ns per frame; the barrier is now one per chunk of
// emission rows, which the staging ring needs anyway.
// Emission rows ([T][row_floats], col 0 = blank) are streamed in chunks of up to 32 frames into a
// 3-deep shared-memory ring by 1-D TMA bulk copies (two chunks in flight: a chunk is consumed in
// well under a microsecond); the next frame's emissions are read into registers before the
// current frame's dependent chain starts. Backpointers are 2-bit step codes (k - bt), one nibble
// per pair per frame, 8 frames per 32-bit word, written coalesced.
// The backtrace is a single warp walking t = T-1..1: lanes hold a 32-pair window of the current
// 8-frame block in registers (next block prefetched); the walker takes the current pair's word by
// shuffle and jumps straight to the next frame whose code is non-zero (count-leading-zeros on the
// masked word), so its cost is one step per block plus one per transition, not one per frame.
// Lane 0 emits first / last+1 at every label-state run boundary (the path is monotone, so each
// label's occupancy is one run).
#include "la_common.cuh"

namespace la {

constexpr int kVitChunkMax = 32;   // frames per TMA chunk (fewer when rows are very wide)
constexpr int kVitStages = 3;
constexpr unsigned long long kNotReady = ~0ull;   // all-ones NaN: neither arithmetic nor an f32->f64 promotion can produce it

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double ring_take(unsigned long long* slot) {
    const uint32_t a = smem_u32(slot);
    unsigned long long v;
    do {
        asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    } while (v == kNotReady);
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(kNotReady) : "memory");   // re-arm
    return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void ring_put(unsigned long long* slot, double v) {
    asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(slot)), "l"(__double_as_longlong(v)) : "memory");
}

// K = pairs per thread, DUMP = parity instrumentation (full fp64 table to global memory; compiled out of
// the production kernels)
template <int K, bool DUMP>
__global__ void __launch_bounds__(1024) viterbi_kernel(const VitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw_launch = blockDim.x >> 5;

    // ---- shared memory carve-up: 3 stages, 3 mbarriers, 2 finals, hand-off ring ----
    const int chunk = p.chunk;
    const int ring = 2 * chunk;                           // slots per warp boundary (power of two)
    const int stage_bytes = chunk * p.row_floats_max * 4;
    float* stage0 = reinterpret_cast<float*>(smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kVitStages * stage_bytes);
    double* fin = reinterpret_cast<double*>(smem + kVitStages * stage_bytes + 32);
    unsigned long long* xchg = reinterpret_cast<unsigned long long*>(smem + kVitStages * stage_bytes + 64);   // [ring][nw_launch]

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
        for (int s = 0; s < kVitStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < ring * nw_launch; i += 32 * nwarps) xchg[i] = kNotReady;
    __syncthreads();

    const int nchunks = (T + chunk - 1) / chunk;
    auto issue = [&](int c) {
        const int rows = min(chunk, T - c * chunk);
        const uint32_t bytes = (uint32_t)rows * wrow * 4;
        const int s = c % kVitStages;
        fence_proxy_async();
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(reinterpret_cast<unsigned char*>(stage0) + s * stage_bytes, E + (int64_t)c * chunk * wrow, bytes, &full[s]);
    };
    if (tid == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }

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
    const bool feeds_right = (lane == 31) && (warp + 1 < nwarps);
    const bool fed_from_left = (lane == 0) && (warp > 0);

    for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[c % kVitStages], (c / kVitStages) & 1);
        if (tid == 0 && c + 2 < nchunks) issue(c + 2);    // stage (c+2)%3 was drained in chunk c-1 (barrier below)
        const float* rows = stage0 + (c % kVitStages) * (stage_bytes / 4);
        const int t0 = c * chunk;
        const int nt = min(chunk, T - t0);
        float fb = rows[0], fl[K];
#pragma unroll
        for (int j = 0; j < K; ++j) fl[j] = rows[ecol[j]];
        for (int tt = 0; tt < nt; ++tt) {
            const int t = t0 + tt;
            const double eb = (double)fb;
            double el[K];
#pragma unroll
            for (int j = 0; j < K; ++j) el[j] = (double)fl[j];
            if (tt + 1 < nt) {                             // next frame's emissions: off the dependent chain
                const float* nr = rows + (tt + 1) * wrow;
                fb = nr[0];
#pragma unroll
                for (int j = 0; j < K; ++j) fl[j] = nr[ecol[j]];
            }
            if (t == 0) {
                // row 0 preset (utils/alignment.py:151-152)
                if (tid == 0) { pb[0] = eb; pl[0] = el[0]; }
                if (feeds_right) ring_put(&xchg[warp], pl[K - 1]);
                if (DUMP) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (pair0 + j <= L) p.dp_dump[2 * (pair0 + j)] = pb[j];
                        if (pair0 + j < L) p.dp_dump[2 * (pair0 + j) + 1] = pl[j];
                    }
                }
                continue;
            }
            double ql = shfl_up_f64(pl[K - 1], 1);
            if (lane == 0) ql = -INFINITY;                 // pair 0: no left neighbour
            if (fed_from_left) ql = ring_take(&xchg[((t - 1) & (ring - 1)) * nw_launch + warp - 1]);
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
            if (feeds_right) ring_put(&xchg[(t & (ring - 1)) * nw_launch + warp], pl[K - 1]);
            if (DUMP) {
                double* drow = p.dp_dump + (int64_t)t * (2 * L + 1);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (pair0 + j <= L) drow[2 * (pair0 + j)] = pb[j];
                    if (pair0 + j < L) drow[2 * (pair0 + j) + 1] = pl[j];
                }
            }
            if (sh == 28 || t == T - 1) {
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
        }
        __syncthreads();                                   // stage fully read before it is refilled; bounds the warps' skew to one chunk
    }

    // ---- end-state pick (utils/alignment.py:157): S-1 iff dp[T-1][S-1] > dp[T-1][S-2] --------
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (pair0 + j == L) fin[0] = pb[j];
        if (pair0 + j == L - 1) fin[1] = pl[j];
    }
    __syncthreads();                                       // also orders the bp stores of all warps before the walker's loads
    if (warp != 0) return;

    int k = (fin[0] > fin[1]) ? 2 * L : 2 * L - 1;
    const double best = (fin[0] > fin[1]) ? fin[0] : fin[1];

    // ---- backtrace: one warp, window of 32 pairs x 8 frames in registers ----------------------
    int32_t* first = p.first + l0;
    int32_t* lastp = p.last_plus1 + l0;
    int visited = 0;
    if ((k & 1) && lane == 0) lastp[k >> 1] = T;
    auto load_win = [&](int tb, int base) -> uint32_t {
        const int pr = base + lane;
        return (tb >= 0 && pr >= 0 && pr < pairs_pad) ? __ldcg(bp + (int64_t)tb * pairs_pad + pr) : 0u;
    };
    int tb = (T - 1) >> 3;
    int base_cur = (k >> 1) - 31;
    uint32_t w_cur = load_win(tb, base_cur);
    while (tb >= 0) {
        const int base_nxt = (k >> 1) - 31;               // covers the <= 16 pairs the next 2 blocks can reach
        const uint32_t w_nxt = load_win(tb - 1, base_nxt);
        const int t_lo = max(1, tb * 8);
        int t = min(T - 1, tb * 8 + 7);
        while (t >= t_lo) {
            // codes of state k over this block: label states use bits 1-2 of each nibble, blank states bit 0
            const uint32_t word = __shfl_sync(0xffffffffu, w_cur, (k >> 1) - base_cur);
            uint32_t m = (k & 1) ? (word & 0x66666666u) : (word & 0x11111111u);
            m &= (0xffffffffu >> (28 - (t & 7) * 4));      // frames above t are already behind the walker
            m &= ~((1u << ((t_lo & 7) * 4)) - 1u);         // frame 0 carries no code (block 0 only)
            if (m == 0u) break;                            // state k stays for the rest of the block
            t = (tb << 3) + ((31 - __clz(m)) >> 2);        // latest frame <= t with a non-zero code
            const uint32_t nib = (word >> ((t & 7) * 4)) & 0xFu;
            const int code = (k & 1) ? (int)(nib >> 1) : (int)(nib & 1u);
            if (k & 1) {                                   // label state k occupied frames t..: onset
                if (lane == 0) first[k >> 1] = t;
                ++visited;
            }
            k -= code;
            if ((k & 1) && lane == 0) lastp[k >> 1] = t;   // new label state ends at frame t-1
            --t;
        }
        w_cur = w_nxt;
        base_cur = base_nxt;
        --tb;
    }
    if (k & 1) {
        if (lane == 0) first[k >> 1] = 0;
        ++visited;
    }
    if (lane == 0) {
        p.status[utt] = (visited == L) ? 0 : 2;          // a missing label state -> ValueError upstream
        p.score[utt] = best;
    }
}

// frames per chunk: a power of two (the hand-off ring is indexed with a mask), at most 32, and small
// enough that three stages of the widest row fit comfortably beside other resident CTAs
int viterbi_chunk_frames(int row_floats_max) {
    const int fit = (36 * 1024) / (row_floats_max * 4);
    int c = kVitChunkMax;
    while (c > 1 && c > fit) c >>= 1;
    return c;
}

size_t viterbi_smem_bytes(int row_floats_max, int chunk, int warps) {
    return (size_t)kVitStages * chunk * row_floats_max * 4 + 64 + (size_t)2 * chunk * warps * 8;
}

template <int K>
static cudaError_t launch_one(const VitParams& p, int threads, int grid, size_t smem, cudaStream_t stream) {
    // the opt-in is a per-function attribute of the loaded module: set once per (function, device), not per launch
    static bool attr_done[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int d = p.dp_dump ? 1 : 0;
    if (dev >= 0 && dev < 64 && !attr_done[d][dev]) {
        cudaError_t e = p.dp_dump
            ? cudaFuncSetAttribute(viterbi_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
            : cudaFuncSetAttribute(viterbi_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_done[d][dev] = true;
    }
    if (p.dp_dump) viterbi_kernel<K, true><<<grid, threads, smem, stream>>>(p);
    else viterbi_kernel<K, false><<<grid, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

// K pairs per lane; one CTA of `warps` warps per utterance
cudaError_t launch_viterbi(const VitParams& p, int K, int warps, cudaStream_t stream) {
    if (p.n_order <= 0) return cudaSuccess;
    const size_t smem = viterbi_smem_bytes(p.row_floats_max, p.chunk, warps);
    switch (K) {
        case 1: return launch_one<1>(p, 32 * warps, p.n_order, smem, stream);
        case 2: return launch_one<2>(p, 32 * warps, p.n_order, smem, stream);
        case 4: return launch_one<4>(p, 32 * warps, p.n_order, smem, stream);
        default: return launch_one<8>(p, 32 * warps, p.n_order, smem, stream);
    }
}

}  // namespace la
