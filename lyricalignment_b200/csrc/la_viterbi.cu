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
// thread that owns K consecutive pairs needs one 64-bit warp shuffle per frame. Two launch
// shapes share the code:
//   * warp-per-utterance (L+1 <= 32K): several utterances per CTA, no block barrier in the loop;
//   * CTA-per-utterance (long-form, up to 32 warps x 32 lanes x K pairs): one double-buffered
//     shared-memory hand-off + one __syncthreads per frame at warp boundaries.
// Emission rows ([T][row_floats], col 0 = blank) are streamed in 16-frame chunks into a
// double-buffered shared-memory window by 1-D TMA bulk copies. Backpointers are 2-bit step
// codes (k - bt), one nibble per pair per frame, 8 frames per 32-bit word, written coalesced.
// The backtrace is a single warp walking t = T-1..1: lanes hold a 32-pair window of the current
// 8-frame block in registers (next block prefetched), the walker reads it by shuffle, and lane 0
// emits first / last+1 at every label-state run boundary (the path is monotone, so each label's
// occupancy is one run).
#include "la_common.cuh"

namespace la {

constexpr int kVitChunkMax = 16;   // frames per TMA chunk (fewer when rows are very wide)

__device__ __forceinline__ double shfl_up_f64(double v, int delta) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return __hiloint2double(hi, lo);
}

// K = pairs per thread, WIDE = all warps of the CTA work on one utterance, DUMP = parity
// instrumentation (full fp64 table to global memory; compiled out of the production kernels)
template <int K, bool WIDE, bool DUMP>
__global__ void __launch_bounds__(WIDE ? 1024 : 128) viterbi_kernel(const VitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int nwarps = blockDim.x >> 5;                        // WIDE: narrowed to this utterance's own need below
    const int group = WIDE ? 0 : warp;                    // smem slice owner
    const int gtid = WIDE ? tid : lane;
    const int slot = WIDE ? blockIdx.x : blockIdx.x * nwarps + warp;

    // ---- shared memory carve-up: per group 2 stages + 2 mbarriers + 2 finals; WIDE: xchg ----
    const int kVitChunk = p.chunk;
    const int stage_bytes = kVitChunk * p.row_floats_max * 4;
    const int group_bytes = 2 * stage_bytes + 32;
    unsigned char* gbase = smem + (size_t)group * group_bytes;
    float* stage0 = reinterpret_cast<float*>(gbase);
    uint64_t* full = reinterpret_cast<uint64_t*>(gbase + 2 * stage_bytes);
    double* fin = reinterpret_cast<double*>(gbase + 2 * stage_bytes + 16);
    double* xchg = reinterpret_cast<double*>(smem + (size_t)(WIDE ? 1 : nwarps) * group_bytes);  // [2][32]

    const bool active = slot < p.n_order;
    if (!WIDE && !active) return;                         // whole warp exits together
    const int utt = active ? p.order[slot] : 0;
    const int T = p.m.t_off[utt + 1] - p.m.t_off[utt];
    const int l0 = p.m.l_off[utt];
    const int L = p.m.l_off[utt + 1] - l0;
    if (L <= 0 || T <= 0) {                               // uniform per group
        if (gtid == 0) {
            p.status[utt] = (L <= 0) ? 1 : 2;
            p.score[utt] = 0.0;
        }
        return;
    }
    const int wrow = p.m.e_row[utt];
    const float* E = p.E + p.m.e_off[utt];
    const int pairs_pad = p.m.bp_pairs[utt];
    uint32_t* bp = p.bp + p.m.bp_off[utt];
    if (WIDE) {
        // the launch is sized for the widest utterance of the bucket; warps this one does not need
        // leave before the first barrier (exited warps do not take part in __syncthreads)
        nwarps = pairs_pad / (32 * K);
        if (warp >= nwarps) return;
    }

    if (gtid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    if (WIDE) __syncthreads(); else __syncwarp();

    const int nchunks = (T + kVitChunk - 1) / kVitChunk;
    auto issue = [&](int c) {
        const int rows = min(kVitChunk, T - c * kVitChunk);
        const uint32_t bytes = (uint32_t)rows * wrow * 4;
        fence_proxy_async();
        mbar_arrive_expect_tx(&full[c & 1], bytes);
        bulk_g2s(reinterpret_cast<unsigned char*>(stage0) + (c & 1) * stage_bytes,
                 E + (int64_t)c * kVitChunk * wrow, bytes, &full[c & 1]);
    };
    if (gtid == 0) issue(0);

    // ---- per-pair constants ------------------------------------------------------------
    const int pair0 = gtid * K;
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

    for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[c & 1], (c >> 1) & 1);
        if (gtid == 0 && c + 1 < nchunks) issue(c + 1);   // stage (c+1)&1 was drained last iteration
        const float* rows = stage0 + (c & 1) * (stage_bytes / 4);
        const int t0 = c * kVitChunk;
        const int nt = min(kVitChunk, T - t0);
        for (int tt = 0; tt < nt; ++tt) {
            const int t = t0 + tt;
            const float* er = rows + tt * wrow;
            if (t == 0) {
                // row 0 preset (utils/alignment.py:151-152)
                if (gtid == 0) { pb[0] = (double)er[0]; pl[0] = (double)er[1]; }
                if (WIDE) {
                    if (lane == 31) xchg[warp] = pl[K - 1];
                    __syncthreads();
                }
                if (DUMP) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (pair0 + j <= L) p.dp_dump[2 * (pair0 + j)] = pb[j];
                        if (pair0 + j < L) p.dp_dump[2 * (pair0 + j) + 1] = pl[j];
                    }
                }
                continue;
            }
            const double eb = (double)er[0];
            double ql = shfl_up_f64(pl[K - 1], 1);
            if (lane == 0) {
                ql = -INFINITY;                            // pair 0: no left neighbour
                if (WIDE && warp > 0) ql = xchg[((t - 1) & 1) * 32 + warp - 1];
            }
            const int sh = (t & 7) * 4;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const double el = (double)er[ecol[j]];
                const double b = pb[j], l = pl[j];
                const double q = ql;
                ql = l;                                    // left neighbour of pair j+1 (old value)
                // blank state 2i (utils/alignment.py:78-82, 92-101)
                const bool b_stay = b > q;
                pb[j] = (b_stay ? b : q) + eb;
                // label state 2i+1 (:84-90, 103-117)
                const bool skip = (q >= b) && (q >= l) && skip_ok[j];
                const bool l_stay = l > b;
                pl[j] = (skip ? q : (l_stay ? l : b)) + el;
                const uint32_t nib = (b_stay ? 0u : 1u) | (skip ? 4u : (l_stay ? 0u : 2u));
                acc[j] |= nib << sh;
            }
            if (WIDE) {
                if (lane == 31) xchg[(t & 1) * 32 + warp] = pl[K - 1];
                __syncthreads();
            }
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
        if (WIDE) __syncthreads(); else __syncwarp();     // stage fully read before it is refilled
    }

    // ---- end-state pick (utils/alignment.py:157): S-1 iff dp[T-1][S-1] > dp[T-1][S-2] --------
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (pair0 + j == L) fin[0] = pb[j];
        if (pair0 + j == L - 1) fin[1] = pl[j];
    }
    if (WIDE) __syncthreads(); else __syncwarp();
    if (WIDE && warp != 0) return;

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
        const int t_hi = min(T - 1, tb * 8 + 7);
        const int t_lo = max(1, tb * 8);
        for (int t = t_hi; t >= t_lo; --t) {
            const uint32_t word = __shfl_sync(0xffffffffu, w_cur, (k >> 1) - base_cur);
            const uint32_t nib = (word >> ((t & 7) * 4)) & 0xFu;
            const int code = (k & 1) ? (int)(nib >> 1) : (int)(nib & 1u);
            if (code) {
                if (k & 1) {                              // label state k occupied frames t..: onset
                    if (lane == 0) first[k >> 1] = t;
                    ++visited;
                }
                k -= code;
                if ((k & 1) && lane == 0) lastp[k >> 1] = t;   // new label state ends at frame t-1
            }
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

int viterbi_chunk_frames(int row_floats_max) {
    int c = (96 * 1024) / (row_floats_max * 4);
    return c < 1 ? 1 : (c > kVitChunkMax ? kVitChunkMax : c);
}

size_t viterbi_smem_bytes(int row_floats_max, int chunk, int groups, int nwarps, bool wide) {
    const size_t group_bytes = 2 * (size_t)chunk * row_floats_max * 4 + 32;
    return group_bytes * groups + (wide ? 2 * 32 * sizeof(double) : 0);
}

template <int K, bool WIDE>
static cudaError_t launch_one(const VitParams& p, int threads, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e;
    if (p.dp_dump) {
        e = cudaFuncSetAttribute(viterbi_kernel<K, WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        viterbi_kernel<K, WIDE, true><<<grid, threads, smem, stream>>>(p);
    } else {
        e = cudaFuncSetAttribute(viterbi_kernel<K, WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        viterbi_kernel<K, WIDE, false><<<grid, threads, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

// K pairs per lane; wide: one CTA of `warps` warps per utterance, else 4 utterances (warps) per CTA
cudaError_t launch_viterbi(const VitParams& p, int K, int warps, bool wide, cudaStream_t stream) {
    if (p.n_order <= 0) return cudaSuccess;
    constexpr int kWarpsPerCta = 4;
    if (!wide) {
        const int grid = (p.n_order + kWarpsPerCta - 1) / kWarpsPerCta;
        const size_t smem = viterbi_smem_bytes(p.row_floats_max, p.chunk, kWarpsPerCta, kWarpsPerCta, false);
        return launch_one<1, false>(p, 32 * kWarpsPerCta, grid, smem, stream);
    }
    const size_t smem = viterbi_smem_bytes(p.row_floats_max, p.chunk, 1, warps, true);
    switch (K) {
        case 1: return launch_one<1, true>(p, 32 * warps, p.n_order, smem, stream);
        case 2: return launch_one<2, true>(p, 32 * warps, p.n_order, smem, stream);
        case 4: return launch_one<4, true>(p, 32 * warps, p.n_order, smem, stream);
        default: return launch_one<8, true>(p, 32 * warps, p.n_order, smem, stream);
    }
}

}  // namespace la
