// la_viterbi_wave.cuh -- K3, the wavefront kernel (included by la_viterbi.cu): utterances of up to 639 pairs.
//
// In the row-synchronous mapping (viterbi_kernel) the loop-carried chain of a frame is DADD -> SHFL -> compare ->
// select -> DADD: the shuffle (and the masking of lane 0 behind it) is a third of it. Here lane i runs i frames
// BEHIND lane 0 (at step tau it works on frame tau - i), so the value it needs from its left neighbour -- that
// lane's label score of frame tau - i - 1 -- was final one whole step earlier: the shuffle is issued a step ahead
// and leaves the chain, which is then compare -> select -> DADD. The price is 31 extra steps per warp.
// * One warp up to 63 pairs (the whole 2 000-clip batch; one pair per lane up to 32 pairs, two from 33 on, chosen
//   per utterance inside ONE launch); beyond that `W` warps of 64 columns each form a pipeline: warp w+1 runs >= 40
//   steps behind warp w and takes the score of w's last pair from a 64-slot shared-memory ring. No sentinels, no
//   barrier, no fence: each warp publishes its step count (a volatile store by the lane that wrote the ring) once
//   per 8-step block, the consumer re-reads it only when the value it remembers is not enough, and a producer more
//   than a ring ahead of its consumer waits the same way.
// * Every warp stages ITS OWN columns of the emission rows, S stages of C rows: a lone warp with ONE bulk copy (TMA)
//   per chunk -- whole rows are contiguous; a warp of a wider utterance with 16-byte asynchronous copies (two
//   256-byte row slices per instruction; per-lane bulk copies serialise into a ~12-instruction loop per row). The
//   first 7 rows of stage 0 are mirrored behind the last stage so an 8-step block never wraps. Lane i reads row
//   (tau - i) at its own column: the lane stride is (K - pitch) floats, odd in units of the access size, so the
//   loads are bank-conflict free. The blank column would be a 4..32-way conflict (the pitch is a multiple of 4): it
//   is fetched separately, 4 bytes per lane and row, into a compact ring with the same slots.
// * Rows -31..0 read as zeros and every state starts at the floor, so the steps a lane runs before its frame 1
//   leave it at exactly -1e7 (= the reference's untouched dp row 0, utils/alignment.py:144-152); lane 0 of warp 0
//   holds the row-0 presets and never executes frame 0. Steps past frame T-1 compute garbage that flows only into
//   later garbage (the dependency runs left to right and forward in time) and into nibbles the walker masks.
// * K = 2 lanes own pairs 2i-1 and 2i (columns 2i, 2i+1 of the row: ONE aligned 8-byte load); pair -1 is a dummy
//   pinned at -inf.
// * Backpointers leave the kernel un-skewed (one funnel shift per 8-step block): word row r holds frames 8r..8r+7.
#pragma once
// (included inside namespace la)

constexpr int kWvMirror = 7;
constexpr int kWvHand = 64;        // hand-off ring slots per warp boundary
constexpr int kWvSlice = 64;       // columns per warp (multi-warp shapes, K = 2)

// Progress words: plain volatile shared-memory accesses. The publisher (lane 31) stores its hand-off values and then
// the step count from ONE thread to ONE memory (shared), which the SM performs in program order; a st.release here
// costs a MEMBAR.ALL.CTA per 8-step block that also waits for the block's backpointer stores to reach L2 (measured:
// the ten-warp long-form trellis ran at 468 cycles per frame with it).
__device__ __forceinline__ uint32_t ld_progress(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void st_progress(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
// per-thread asynchronous copies (LDGSTS): 16 bytes for row slices, 4 bytes for the strided blank column
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(dst), "l"(src), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
                 ::"r"(dst), "l"(src), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// K pairs per lane, C rows per chunk, S stages, MULTI: more than one warp per utterance (K = 2, 64-column slices)
template <int K, int C, int S, bool MULTI, bool DUMP>
__device__ __forceinline__ void wave_run(const VitParams& p, unsigned char* smem, int utt, int T, int l0, int L) {
    constexpr int SH = (K == 1) ? 0 : 1;                  // column of pair i = i + SH
    constexpr int LOGK = (K == 1) ? 0 : 1;
    constexpr int R = C * S;                              // staged rows (+ 7 mirrored)
    constexpr int WIN = (31 + C - 1) / C;                 // chunks behind the current one that a warp still reads
    constexpr int SPB = C / 8;                            // 8-step blocks per chunk
    constexpr int AHEAD = S - WIN - 1;                    // chunks in flight beyond the current one
    static_assert(C % 8 == 0 && C <= 32 && AHEAD >= 1, "chunk/stage shape");
    static_assert(!MULTI || K == 2, "multi-warp shapes use two pairs per lane");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wrow = p.m.e_row[utt];
    const float* E = p.E + p.m.e_off[utt];
    const int pairs_pad = p.m.bp_pairs[utt];
    uint32_t* bp = p.bp + p.m.bp_off[utt];
    const int col_last = L + SH;                          // column of pair L
    const int w_last = MULTI ? (col_last >> 6) : 0;       // warp that owns it
    if (warp > w_last) return;                            // the launch is sized for the widest utterance of its bucket
    const int nw_launch = blockDim.x >> 5;

    // ---- shared memory: per warp [R + 7 rows][pitch], blank ring [R + 8], S barriers; then progress words + hand-off rings
    const int pitch = MULTI ? kWvSlice : wrow;            // floats between staged rows
    const int pitch_alloc = MULTI ? kWvSlice : p.row_floats_max;
    const size_t rows_bytes = max((size_t)(R + kWvMirror) * pitch_alloc * 4 + 256, kBtSmemBytes);   // the walker reuses warp 0's
    const size_t warp_bytes = rows_bytes + (R + 8) * 4 + 64;
    float* rows = reinterpret_cast<float*>(smem + warp * warp_bytes);
    float* bl = reinterpret_cast<float*>(smem + warp * warp_bytes + rows_bytes);              // compact blank column, same slots
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + warp * warp_bytes + rows_bytes + (R + 8) * 4);
    unsigned char* shared0 = smem + nw_launch * warp_bytes;
    uint32_t* prog = reinterpret_cast<uint32_t*>(shared0);                                    // [nw_launch] (+ fin[2] at +64)
    double* hand = reinterpret_cast<double*>(shared0 + 128);                                  // [nw_launch][kWvHand + 8]

    if (!MULTI && lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    if (MULTI && lane == 0) prog[warp] = 0u;
    // frames -C*WIN..-1 live in the last WIN stages until real chunks replace them
    for (int i = lane; i < WIN * C * pitch; i += 32) rows[(R - WIN * C) * pitch + i] = 0.f;
    for (int i = lane; i < WIN * C; i += 32) bl[R - WIN * C + i] = 0.f;
    if (!MULTI) fence_proxy_async();                      // the zeroed rows are overwritten by bulk copies later on
    __syncwarp();
    const int nchunks = (T + C - 1) / C;
    const int col_base = MULTI ? warp * kWvSlice : 0;
    const int slice_cols = MULTI ? min(kWvSlice, wrow - col_base) : wrow;
    const uint32_t rows_u32 = smem_u32(rows), bl_u32 = smem_u32(bl);
    // whole warp: chunk c into stage st (+ the mirror rows), one commit group per call (empty past the last chunk)
    auto issue = [&](int c, int st) {
        if (c < nchunks) {
            const int nr = min(C, T - c * C);
            const int nm = st == 0 ? min(kWvMirror, nr) : 0;
            const float* src = E + (int64_t)c * C * wrow;
            if (MULTI) {                                  // 16 bytes per lane: two 256-byte row slices per instruction
                const int c4 = (lane & 15) * 4, rr = lane >> 4;
                const float* s16 = src + col_base + c4 + (int64_t)rr * wrow;
                const uint32_t d16 = rows_u32 + (uint32_t)((st * C + rr) * kWvSlice + c4) * 4u;
#pragma unroll
                for (int j = 0; j < C / 2; ++j)
                    cp_async16(d16 + j * 2 * kWvSlice * 4, s16 + (int64_t)j * 2 * wrow, (2 * j + rr < nr) && c4 < slice_cols);
                if (st == 0) {
                    const uint32_t m16 = rows_u32 + (uint32_t)((R + rr) * kWvSlice + c4) * 4u;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        cp_async16(m16 + j * 2 * kWvSlice * 4, s16 + (int64_t)j * 2 * wrow, (2 * j + rr < nm) && c4 < slice_cols);
                }
            } else if (lane == 0) {                       // whole rows are contiguous: one bulk copy
                const uint32_t row_bytes = (uint32_t)wrow * 4u;
                mbar_arrive_expect_tx(&full[st], (uint32_t)(nr + nm) * row_bytes);
                bulk_g2s(rows + st * C * pitch, src, (uint32_t)nr * row_bytes, &full[st]);
                if (nm) bulk_g2s(rows + R * pitch, src, (uint32_t)nm * row_bytes, &full[st]);
            }
            // the blank column (column 0 of the full row), 4 bytes per lane
            cp_async4(bl_u32 + (uint32_t)(st * C + lane) * 4u, src + (int64_t)lane * wrow, lane < nr);
            if (st == 0) cp_async4(bl_u32 + (uint32_t)(R + lane) * 4u, src + (int64_t)lane * wrow, lane < nm);
        }
        cp_async_commit();
    };
    for (int c = 0; c <= AHEAD; ++c) issue(c, c);
    if (MULTI) __syncthreads();                           // prog[] zeroed before anyone polls it

    // ---- per-lane constants ----------------------------------------------------------------
    const int col0 = col_base + K * lane;                 // first column of the lane; pair = column - SH
    const int pair0 = col0 - SH;
    bool skip_ok[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int i = pair0 + j;
        skip_ok[j] = (i >= 1 && i < L) ? (p.m.labels[l0 + i] != p.m.labels[l0 + i - 1]) : false;
    }
    const bool force0 = threadIdx.x == 0;                 // K = 1: pair 0 has no left neighbour; K = 2: the dummy pair
    // emission column of the lane inside its staged rows (pair i reads column 1 + i). Lanes past the utterance's last
    // column keep the SAME address pattern and read whatever lies there (the next row, stale bytes: their pairs do not
    // exist and feed nothing real): parking them all on column 0 put them on one bank -- a 32-way conflict on every load
    // of the last warp of a wide utterance, which then set the pace of the whole pipeline (measured: 205 vs 153 cycles
    // per step with 2 warps). The row area has 64 floats of slack for the overrun.
    const int ecol = (K == 1) ? 1 + pair0 : K * lane;
    const bool has_left = MULTI && warp > 0, has_right = MULTI && warp < w_last;
    uint32_t lane0 = lane == 0, lane31 = lane == 31;
    const uint32_t prog_mine = smem_u32(prog + warp), prog_left = smem_u32(prog + max(warp, 1) - 1),
                   prog_right = smem_u32(prog + min(warp + 1, nw_launch - 1));
    uint32_t hand_put = smem_u32(hand + warp * (kWvHand + 8)), hand_take = smem_u32(hand + (max(warp, 1) - 1) * (kWvHand + 8));
    asm volatile("" : "+r"(lane0), "+r"(lane31), "+r"(hand_put), "+r"(hand_take));   // opaque: keep them in registers

    cp_async_wait<AHEAD>();
    if (!MULTI) mbar_wait(&full[0], 0);
    __syncwarp();
    float e00 = 0.f, e01 = 0.f;
    if (warp == 0) { e00 = bl[0]; e01 = rows[1]; }
    __syncwarp();
    for (int i = lane; i < pitch; i += 32) { rows[i] = 0.f; rows[R * pitch + i] = 0.f; }   // row 0 (and its mirror) as zeros
    if (lane == 0) { bl[0] = 0.f; bl[R] = 0.f; }
    if (!MULTI) fence_proxy_async();                      // (the only generic-proxy WRITES to the stages; refills after
    __syncwarp();                                         // this point follow generic READS and need no proxy fence)

    double b[K], l[K];
    uint32_t acc[K], prev[K];                             // codes of the current / the previous 8-step block
#pragma unroll
    for (int j = 0; j < K; ++j) { b[j] = kFloor; l[j] = kFloor; acc[j] = 0u; prev[j] = 0u; }
    if (threadIdx.x == 0) {                               // row 0 presets (utils/alignment.py:151-152)
        if (K == 1) { b[0] = (double)e00; l[0] = (double)e01; }
        else { b[0] = -INFINITY; l[0] = -INFINITY; b[K - 1] = (double)e00; l[K - 1] = (double)e01; }
    }
    double q = kFloor;                                    // left neighbour's label score, one frame back

    const int i_last = (col_last >> LOGK) & 31;           // lane of pair L (in warp w_last)
    const int nsteps = T + (warp == w_last ? i_last : 31);   // steps 1 .. nsteps-1; the last lane that matters ends on frame T-1

    // one pair-frame: the reference's comparisons (utils/alignment.py:78-117) with (q >= b && q >= l) folded into
    // q >= max(b, l) -- identical for the finite / -inf values that occur
#define LA_WAVE_CELL(J, EB, EL, SHIFT, ACC)                                               \
    {                                                                                       \
        const double bj = b[J], lj = l[J];                                                  \
        const bool P_b = (J == 0) ? ((bj > qq) || force0) : (bj > qq);                      \
        const bool P_l = lj > bj;                                                           \
        const double alt = P_l ? lj : bj;                                                   \
        const bool P_s = (qq >= alt) && skip_ok[J];                                         \
        b[J] = (P_b ? bj : qq) + (EB);                                                      \
        l[J] = (P_s ? qq : alt) + (EL);                                                     \
        ACC |= ((P_b ? 0u : 1u) | (P_s ? 4u : (P_l ? 0u : 2u))) << (SHIFT);                 \
        qq = lj;                                                                            \
    }
    // Hand-off ring: the value of frame f sits in slot f % 64. Lane 0 of a warp with a left neighbour overwrites the
    // q it got from the shuffle; lane 31 of a warp with a right neighbour stores its last pair's label score.
    auto take = [&](double& qn, uint32_t addr) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.volatile.shared.f64 %0, [%1];\n\t}"
                     : "+d"(qn) : "r"(addr), "r"(lane0));
    };
    auto put = [&](uint32_t addr, double v) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.volatile.shared.f64 [%0], %1;\n\t}"
                     ::"r"(addr), "d"(v), "r"(lane31));
    };
    auto hand_slot = [](int f) { return (uint32_t)(f & (kWvHand - 1)) * 8u; };

    // Backpointers leave the kernel UN-skewed: word row r of a column holds frames 8r .. 8r+7. An 8-step block of lane
    // i covers frames tau0 - i .. tau0 - i + 7, which straddle two such words unless i % 8 == 0: the word that the
    // block completes is the top `o` nibbles of the previous block's codes followed by the bottom 8 - o of this one's
    // (o = -i mod 8) -- one funnel shift -- and it belongs to row tau0/8 - ceil(i/8).
    const int fs = 32 - 4 * ((-lane) & 7);
    const int nrows = (T + 7) >> 3;
    uint32_t* const bp_lane = bp + col0 - (int64_t)((lane + 7) >> 3) * pairs_pad;
    auto emit = [&](int tau0, const uint32_t (&cur)[K]) {
        const int row = (tau0 >> 3) - ((lane + 7) >> 3);
        uint32_t w[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { w[j] = __funnelshift_rc(prev[j], cur[j], fs); prev[j] = cur[j]; }
        // a predicated store, not a branch: the rows differ between lanes, and a divergent branch here costs the
        // lone warp ~150 cycles per block (reconvergence before the next shuffle)
        const uint32_t ok = (unsigned)row < (unsigned)nrows;
        uint32_t* dst = bp_lane + (int64_t)(tau0 >> 3) * pairs_pad;
        if (K == 2)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.v2.b32 [%0], {%1, %2};\n\t}"
                         ::"l"(dst), "r"(w[0]), "r"(w[K - 1]), "r"(ok) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.b32 [%0], %1;\n\t}"
                         ::"l"(dst), "r"(w[0]), "r"(ok) : "memory");
    };

    // any single step (the first block, the tail, DUMP)
    auto step = [&](int tau, const float* pe, const float* pb) {
        double qn = shfl_up_f64(l[K - 1], 1);
        if (has_left) take(qn, hand_take + hand_slot(tau));
        const double eb = (double)pb[0];
        double qq = q;
        const int sh4 = (tau & 7) * 4;
#pragma unroll
        for (int j = 0; j < K; ++j) LA_WAVE_CELL(j, eb, (double)pe[j], sh4, acc[j])
        q = qn;
        if (has_right) {
            const int f = tau - 31;
            put(hand_put + hand_slot(f), l[K - 1]);
        }
        if (DUMP) {
            const int t = tau - lane;
            if (t >= 1 && t < T) {
                double* drow = p.dp_dump + (int64_t)t * (2 * L + 1);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int pr = pair0 + j;
                    if (pr >= 0 && pr <= L) drow[2 * pr] = b[j];
                    if (pr >= 0 && pr < L) drow[2 * pr + 1] = l[j];
                }
            }
        }
        if ((tau & 7) == 7 || tau == nsteps - 1) {
            emit(tau & ~7, acc);
#pragma unroll
            for (int j = 0; j < K; ++j) acc[j] = 0u;
        }
    };
    // eight whole steps tau0 .. tau0+7 (tau0 a multiple of 8): emissions in registers up front, immediates everywhere
    auto fast8 = [&](int tau0, const float* pe, const float* pb, auto HL, auto HR) {
        constexpr bool kLeft = decltype(HL)::value, kRight = decltype(HR)::value;
        float ebf[8], elf[8][K];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ebf[i] = pb[i];
            if (K == 2) {
                const float2 v = *reinterpret_cast<const float2*>(pe + i * pitch);
                elf[i][0] = v.x; elf[i][K - 1] = v.y;
            } else {
                elf[i][0] = pe[i * pitch];
            }
        }
        // the consumer reads frames tau0 .. tau0+7 = slots s .. s+7 (s a multiple of 8); the producer writes frames
        // tau0-31 .. tau0-24 = slots s'+1 .. s'+8 (s' = (tau0-32) % 64): when s'+8 is 64 (a scratch slot) the value
        // belongs in slot 0 and is stored again after the loop
        const uint32_t take8 = hand_take + hand_slot(tau0);
        const uint32_t put8 = hand_put + hand_slot(tau0 - 32) + 8u;
        uint32_t a[K];
#pragma unroll
        for (int j = 0; j < K; ++j) a[j] = 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double qn = shfl_up_f64(l[K - 1], 1);
            if (kLeft) take(qn, take8 + 8u * i);
            const double eb = (double)ebf[i];
            double qq = q;
#pragma unroll
            for (int j = 0; j < K; ++j) LA_WAVE_CELL(j, eb, (double)elf[i][j], 4 * i, a[j])
            q = qn;
            if (kRight) put(put8 + 8u * i, l[K - 1]);
        }
        if (kRight && ((tau0 - 32) & (kWvHand - 1)) == kWvHand - 8) put(hand_put, l[K - 1]);   // frame tau0-24 -> slot 0
        emit(tau0, a);
    };

    const long long c_fwd0 = clock64();
    int slot = lane ? R - lane : 0;                       // staged row of frame tau0 - lane, tau0 = 0
    int st = 0, ph = 0;                                   // stage / parity of the chunk the current block starts
    int st_fill = AHEAD + 1;                              // stage that receives the next refill (that of chunk c - WIN - 1)
    uint32_t seen_left = 0u, seen_right = 0u;             // neighbours' progress words as last read
    for (int sb = 0; sb * 8 < nsteps; ++sb) {
        const int tau0 = sb * 8;
        if (sb % SPB == 0 && sb) {                        // ---- chunk boundary: refill the stage that died, wait for chunk c
            const int c = sb / SPB;
            __syncwarp();                                 // every lane is past chunk c - WIN - 1
            issue(c + AHEAD, st_fill);
            if (++st_fill == S) st_fill = 0;
            if (++st == S) { st = 0; ph ^= 1; }
            cp_async_wait<AHEAD>();
            if (!MULTI && c < nchunks) mbar_wait(&full[st], ph);
            __syncwarp();
        }
        if (MULTI) {                                      // ---- flow control against the neighbours, once per block
            // the last value seen is remembered: a neighbour that was already far enough costs no shared-memory round trip
            if (has_left) {
                const uint32_t need = (uint32_t)min(tau0 + 8 + 31, T + 31);   // slots up to tau0 + 7 written
                while (seen_left < need) seen_left = ld_progress(prog_left);
            }
            if (has_right) {
                const int need = tau0 - 23 - kWvHand;     // slots up to tau0 - 24 - kWvHand consumed
                while ((int)seen_right < need) seen_right = ld_progress(prog_right);
            }
        }
        const float* pe = rows + slot * pitch + ecol;
        const float* pb = bl + slot;
        if (!DUMP && tau0 >= 8 && tau0 + 8 <= nsteps) {
            if (!MULTI) fast8(tau0, pe, pb, std::false_type{}, std::false_type{});
            else if (has_left && has_right) fast8(tau0, pe, pb, std::true_type{}, std::true_type{});
            else if (has_left) fast8(tau0, pe, pb, std::true_type{}, std::false_type{});
            else if (has_right) fast8(tau0, pe, pb, std::false_type{}, std::true_type{});
            else fast8(tau0, pe, pb, std::false_type{}, std::false_type{});
        } else {
            const int hi = min(tau0 + 8, nsteps);
            for (int tau = max(tau0, 1); tau < hi; ++tau) step(tau, pe + (tau - tau0) * pitch, pb + (tau - tau0));
        }
        if (MULTI) {
            __syncwarp();
            if (lane31) st_progress(prog_mine, (uint32_t)min(tau0 + 8, nsteps));
        }
        slot += 8;
        if (slot >= R) slot -= R;
    }
#undef LA_WAVE_CELL
    {                                                     // the codes still waiting for the top of their word
        const uint32_t zero[K] = {};
        emit(((nsteps - 1) & ~7) + 8, zero);
    }

    // ---- end-state pick (utils/alignment.py:157): S-1 iff dp[T-1][S-1] > dp[T-1][S-2]. Lane i_last of the last warp
    // stopped on frame T-1; pair L-1's label score of that frame is its own other slot or the q it would use next.
    double* fin = reinterpret_cast<double*>(shared0 + 64);                                    // [2] inside the prog block
    if (warp == w_last) {
        const int jL = col_last & (K - 1);
        const double f0 = (K == 2 && jL) ? b[K - 1] : b[0];
        const double f1 = (K == 2 && jL) ? l[0] : q;
        if (lane == i_last) { fin[0] = f0; fin[1] = f1; }
    }
    __threadfence_block();
    if (MULTI) __syncthreads(); else __syncwarp();        // every warp done; their bp stores are ordered before the walker's loads
    if (warp != 0) return;
    const long long c_fwd1 = clock64();
    const int k = (fin[0] > fin[1]) ? 2 * L : 2 * L - 1;
    const double best = (fin[0] > fin[1]) ? fin[0] : fin[1];
    const int visited = backtrace_walk<SH>(bp, pairs_pad, T, k, lane, p.first + l0, p.last_plus1 + l0,
                                                       reinterpret_cast<uint32_t*>(rows));   // warp 0's stages: the DP is done with them
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

// Stage shapes. A lone warp with the SM to itself: 4 stages of 32 rows (64 steps of look-ahead, a chunk boundary
// every 32 steps). A batch (more utterances than two per SM): 4 stages of 16 rows -- 14 KB per CTA at 48-float rows
// instead of 27 KB, so 16 instead of 8 one-warp CTAs share an SM and hide each other's latencies (measured on the
// 2 000-clip batch: issue slots 54 % busy at 8 CTAs per SM). Multi-warp shapes: 4 stages of 16 rows (18 KB per warp).
// Up to four warps per CTA the multi-warp shapes use the lone warp's 4 x 32 rows (35 KB per warp).
constexpr int kWv1C = 32, kWv1S = 4, kWvBC = 16, kWvBS = 4, kWvMC = 16, kWvMS = 4;

template <bool MULTI, bool DUMP, int MAXT, int MC = kWvMC, int MS = kWvMS>
__global__ void __launch_bounds__(MAXT) viterbi_wave_kernel(const VitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int utt = p.order[blockIdx.x];
    const int T = p.m.t_off[utt + 1] - p.m.t_off[utt];
    const int l0 = p.m.l_off[utt];
    const int L = p.m.l_off[utt + 1] - l0;
    if (L <= 0 || T <= 0) {
        if (threadIdx.x == 0) {
            p.status[utt] = (L <= 0) ? 1 : 2;
            p.score[utt] = 0.0;
        }
        return;
    }
    if (MULTI) wave_run<2, MC, MS, true, DUMP>(p, smem, utt, T, l0, L);
    else if (p.chunk == kWvBC) {
        if (L + 1 <= 32) wave_run<1, kWvBC, kWvBS, false, DUMP>(p, smem, utt, T, l0, L);
        else wave_run<2, kWvBC, kWvBS, false, DUMP>(p, smem, utt, T, l0, L);
    } else {
        if (L + 1 <= 32) wave_run<1, kWv1C, kWv1S, false, DUMP>(p, smem, utt, T, l0, L);
        else wave_run<2, kWv1C, kWv1S, false, DUMP>(p, smem, utt, T, l0, L);
    }
}

static size_t viterbi_wave_smem_bytes(int row_floats_max, int warps, int chunk) {
    const bool multi = warps > 1;
    const int R = multi ? (warps <= 4 ? kWv1C * kWv1S : kWvMC * kWvMS) : (chunk == kWvBC ? kWvBC * kWvBS : kWv1C * kWv1S);
    const size_t warp_bytes = std::max((size_t)(R + kWvMirror) * (multi ? kWvSlice : row_floats_max) * 4 + 256, kBtSmemBytes) + (R + 8) * 4 + 64;
    return warps * warp_bytes + 128 + (multi ? (size_t)warps * (kWvHand + 8) * 8 : 0);
}

template <bool MULTI, int MAXT, int MC = kWvMC, int MS = kWvMS>
static cudaError_t launch_wave(const VitParams& p, int threads, size_t smem, cudaStream_t stream) {
    static bool attr_done[2][64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int d = p.dp_dump ? 1 : 0;
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !attr_done[d][dev]) {
        cudaError_t e = p.dp_dump
            ? cudaFuncSetAttribute(viterbi_wave_kernel<MULTI, true, MAXT, MC, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)
            : cudaFuncSetAttribute(viterbi_wave_kernel<MULTI, false, MAXT, MC, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return e;
        attr_done[d][dev] = true;
    }
    if (p.dp_dump) viterbi_wave_kernel<MULTI, true, MAXT, MC, MS><<<p.n_order, threads, smem, stream>>>(p);
    else viterbi_wave_kernel<MULTI, false, MAXT, MC, MS><<<p.n_order, threads, smem, stream>>>(p);
    return cudaGetLastError();
}

// one CTA of `warps` warps per utterance (1: up to 63 pairs; else 64 columns per warp, up to 10 warps)
cudaError_t launch_viterbi_wave(const VitParams& p_in, int warps, cudaStream_t stream) {
    if (p_in.n_order <= 0) return cudaSuccess;
    VitParams p = p_in;
    static const int trace = [] { const char* e = getenv("LA_VIT_TRACE"); return e ? atoi(e) : 0; }();
    p.trace = trace;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    static const int force = [] { const char* e = getenv("LA_WAVE_CHUNK"); return e ? atoi(e) : 0; }();   // A/B knob
    p.chunk = force ? force : (p.n_order > 2 * sms ? kWvBC : kWv1C);
    const size_t smem = viterbi_wave_smem_bytes(p.row_floats_max, warps, p.chunk);
    if (warps == 1) return launch_wave<false, 32>(p, 32, smem, stream);
    if (warps <= 4) return launch_wave<true, 128, kWv1C, kWv1S>(p, 32 * warps, smem, stream);
    return launch_wave<true, 320>(p, 32 * warps, smem, stream);
}

