// la_emit.cu -- K2: fused log-softmax + label-column gather over the [sum T][V] fp32 logits.
//
// Replaces the reference's emission math (utils/alignment.py:123-134 for CTC-trained models,
// :14-20 for CE-trained ones): 5-6 eager torch passes over B*T*V*4 bytes on the CPU become ONE
// streaming pass that never materialises the [T][V] log-softmax and writes only the compact
// emission rows the DP needs (blank + L label columns per frame).
//
// Shape of the kernel (HBM-bound, 84.5 KB per frame at V = 21129):
//   * persistent CTAs, 3 per SM, each owning a contiguous range of frame rows;
//   * one producer warp streams every row through a 4-stage x 16 KB shared-memory ring with
//     1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); rows are only 4-byte aligned
//     (V is odd), so the copy covers the 16-byte-aligned interior and <= 3 tail floats are read
//     directly;
//   * 8 consumer warps pull float4s out of the ring, release the stage immediately, and keep a
//     per-thread online (max, sum-exp) pair -- one pass over shared memory, FFMA + MUFU.EX2 + FADD
//     per logit (the normaliser only; see ex2_approx);
//   * per row: block reduce of the (max, sum) pairs, then L+1 threads gather the label / silence
//     logits (L2-hot, the row has just streamed through) and apply the reference's exact
//     formulas in fp32: (z - max) - log(sum); naive 1/(1+exp(-z)) sigmoid; log(1 - s); add;
//     clip at -1000 AFTER the add.
#include <cstdlib>

#include "la_common.cuh"

namespace la {

constexpr int kEmitConsumers = 256;
constexpr int kEmitThreads = kEmitConsumers + 32;
constexpr int kEmitStageBytes = 16384;
constexpr int kF4PerThread = kEmitStageBytes / 16 / kEmitConsumers;   // 4

struct RowGeom {
    int64_t a_start;   // byte offset (from the logits base) of the aligned interior
    int nbytes;        // interior bytes, multiple of 16
    int lead;          // floats between a_start and the row start (0..3)
    int nchunks;
    int cb;            // bytes per chunk, multiple of 16
    int64_t tail_start;
    int ntail;
};

__device__ __forceinline__ RowGeom row_geom(int64_t row, int64_t ld, int V) {
    RowGeom g;
    const int64_t rs = row * ld * 4;
    const int64_t re = rs + (int64_t)V * 4;
    g.a_start = rs & ~int64_t(15);
    const int64_t a_end = re & ~int64_t(15);
    g.nbytes = a_end > g.a_start ? (int)(a_end - g.a_start) : 0;
    g.lead = (int)((rs - g.a_start) >> 2);
    g.nchunks = (g.nbytes + kEmitStageBytes - 1) / kEmitStageBytes;
    g.cb = g.nchunks ? ((((g.nbytes + g.nchunks - 1) / g.nchunks) + 15) & ~15) : 0;
    g.tail_start = g.nbytes ? a_end : rs;
    g.ntail = (int)((re - g.tail_start) >> 2);
    return g;
}

// 2^x on the SFU (MUFU.EX2, max rel. error 2^-22). Used ONLY for the sum-exp normaliser, where
// 21127 terms are accumulated and the result goes through a log: the LSE moves by < 3e-7, far
// below the ulp-level differences between libm/Sleef/CUDA expf that the reference's own chain
// already has. The label-column epilogue keeps the exact fp32 formulas with full-precision libm.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// STAGES x 16 KB ring per CTA, CTAS resident CTAs per SM (6 x 2 and 4 x 3 both keep 192 KB in flight per SM)
template <int MODE, int kEmitStages, int CTAS>
__global__ void __launch_bounds__(kEmitThreads, CTAS) emit_kernel(const EmitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kEmitStages * kEmitStageBytes);
    uint64_t* empty = full + kEmitStages;
    float2* red = reinterpret_cast<float2*>(empty + kEmitStages);   // [2][8]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int V = p.m.V;
    const int64_t r0 = (int64_t)blockIdx.x * p.n_rows / gridDim.x;
    const int64_t r1 = (int64_t)(blockIdx.x + 1) * p.n_rows / gridDim.x;
    const unsigned char* gbase = reinterpret_cast<const unsigned char*>(p.logits);

    if (tid == 0) {
        for (int s = 0; s < kEmitStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kEmitConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kEmitConsumers / 32) {
        // ===================== producer warp: one elected lane drives the TMA ring ==========
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();   // every logit is read exactly once
            uint32_t it = 0;
            for (int64_t row = r0; row < r1; ++row) {
                const RowGeom g = row_geom(row, p.ld, V);
                for (int c = 0; c < g.nchunks; ++c, ++it) {
                    const int stage = it % kEmitStages;
                    const uint32_t ph = (it / kEmitStages) & 1;
                    const int off = c * g.cb;
                    const int sz = min(g.cb, g.nbytes - off);
                    mbar_wait(&empty[stage], ph ^ 1);
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)sz);
                    if (p.l2_hint) bulk_g2s_hint(ring + stage * kEmitStageBytes, gbase + g.a_start + off, (uint32_t)sz, &full[stage], pol);
                    else bulk_g2s(ring + stage * kEmitStageBytes, gbase + g.a_start + off, (uint32_t)sz, &full[stage]);
                }
            }
        }
        return;
    }

    // ========================= consumer warps ==============================================
    const int lo = (MODE == 0) ? 1 : 0;          // softmax column range [lo, hi]
    const int hi = (MODE == 0) ? V - 2 : V - 1;
    uint32_t it = 0;
    int utt = -1;
    int64_t utt_r0 = 0, utt_r1 = 0;
    int L = 0, wrow = 0;
    const int32_t* lab = nullptr;
    float* Eutt = nullptr;

    for (int64_t row = r0; row < r1; ++row) {
        const int64_t grow = row + p.row0;                  // batch row (utterance lookup, output row)
        if (grow >= utt_r1 || utt < 0) {
            // binary search: last u with t_off[u] <= row
            int a = 0, b = p.m.n_utt;
            while (b - a > 1) {
                const int mid = (a + b) >> 1;
                if (__ldg(&p.m.t_off[mid]) <= grow) a = mid; else b = mid;
            }
            utt = a;
            utt_r0 = __ldg(&p.m.t_off[utt]);
            utt_r1 = __ldg(&p.m.t_off[utt + 1]);
            const int l0 = __ldg(&p.m.l_off[utt]);
            L = __ldg(&p.m.l_off[utt + 1]) - l0;
            lab = p.m.labels + l0;
            wrow = __ldg(&p.m.e_row[utt]);
            Eutt = p.E + __ldg(&p.m.e_off[utt]);
        }
        const RowGeom g = row_geom(row, p.ld, V);
        const float* rowp = p.logits + row * p.ld;

        float m = -INFINITY;
        float2 s2 = make_float2(0.f, 0.f);                  // two interleaved partial sums (packed fp32x2 math)
        for (int c = 0; c < g.nchunks; ++c, ++it) {
            const int stage = it % kEmitStages;
            const uint32_t ph = (it / kEmitStages) & 1;
            const int sz = min(g.cb, g.nbytes - c * g.cb);
            const int nf4 = sz >> 4;
            mbar_wait(&full[stage], ph);
            const float4* src = reinterpret_cast<const float4*>(ring + stage * kEmitStageBytes);
            float4 v[kF4PerThread];
#pragma unroll
            for (int k = 0; k < kF4PerThread; ++k) {
                const int idx = tid + k * kEmitConsumers;
                v[k] = (idx < nf4) ? src[idx] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);     // stage is free as soon as it is in registers

            const int cbase = ((c * g.cb) >> 2) - g.lead;   // column of float 0 of this chunk
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < kF4PerThread; ++k) {
                const int col0 = cbase + (tid + k * kEmitConsumers) * 4;
                if (col0 < lo || col0 + 3 > hi) {           // row edges only: mask the strangers
                    if (col0 < lo || col0 > hi) v[k].x = -INFINITY;
                    if (col0 + 1 < lo || col0 + 1 > hi) v[k].y = -INFINITY;
                    if (col0 + 2 < lo || col0 + 2 > hi) v[k].z = -INFINITY;
                    if (col0 + 3 < lo || col0 + 3 > hi) v[k].w = -INFINITY;
                }
                mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
            }
            if (mx > -INFINITY) {
                const float mn = fmaxf(m, mx);
                const float sc = ex2_approx((m - mn) * kLog2e);
                const float2 l2 = make_float2(kLog2e, kLog2e), nm = make_float2(-mn * kLog2e, -mn * kLog2e);
                float2 acc = make_float2(s2.x * sc, s2.y * sc);
#pragma unroll
                for (int k = 0; k < kF4PerThread; ++k) {      // per 2 logits: 1 FFMA2 + 2 MUFU.EX2 + 1 FADD2
                    const float2 a = __ffma2_rn(make_float2(v[k].x, v[k].y), l2, nm);
                    const float2 b = __ffma2_rn(make_float2(v[k].z, v[k].w), l2, nm);
                    acc = __fadd2_rn(acc, make_float2(ex2_approx(a.x), ex2_approx(a.y)));
                    acc = __fadd2_rn(acc, make_float2(ex2_approx(b.x), ex2_approx(b.y)));
                }
                s2 = acc;
                m = mn;
            }
        }
        float s = s2.x + s2.y;
        if (tid < g.ntail) {                                // <= 3 floats past the aligned interior
            const int col = (int)((g.tail_start - row * p.ld * 4) >> 2) + tid;
            if (col >= lo && col <= hi) {
                const float x = __ldg(reinterpret_cast<const float*>(gbase + g.tail_start) + tid);
                const float mn = fmaxf(m, x);
                s = s * ex2_approx((m - mn) * kLog2e) + ex2_approx((x - mn) * kLog2e);
                m = mn;
            }
        }
        // issue the gather loads early; they are consumed after the block reduce
        const bool ctc = (MODE == 0);
        const float zsil = ctc ? __ldg(rowp + (V - 1)) : __ldg(rowp);
        float zl = 0.f;
        if (tid < L) zl = __ldg(rowp + __ldg(lab + tid));

        // ---- block reduce of (m, s) over the 8 consumer warps ----------------------------
        float wm = m;
#pragma unroll
        for (int o = 16; o; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        float ws = (m > -INFINITY) ? s * ex2_approx((m - wm) * kLog2e) : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
        float2* rr = red + (row & 1) * (kEmitConsumers / 32);
        if (lane == 0) rr[warp] = make_float2(wm, ws);
        named_bar_sync(1, kEmitConsumers);
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kEmitConsumers / 32; ++w) M = fmaxf(M, rr[w].x);
        float S = 0.f;
#pragma unroll
        for (int w = 0; w < kEmitConsumers / 32; ++w) {
            const float2 q = rr[w];
            S += (q.x > -INFINITY) ? q.y * ex2_approx((q.x - M) * kLog2e) : 0.f;
        }
        const float logS = logf(S);

        // ---- epilogue: the reference's formulas, fp32, same operation order ----------------
        float* Erow = Eutt + (grow - utt_r0) * (int64_t)wrow;
        float add = 0.f, blank;
        if (ctc) {
            const float sg = 1.0f / (1.0f + expf(-zsil));    // F.sigmoid            (:125)
            add = logf(1.0f - sg);                           // log(1 - s)           (:126,129)
            blank = fmaxf(logf(sg), kClip);                  // clip(log s, -1000)   (:128,134)
        } else {
            blank = fmaxf((zsil - M) - logS, kClip);         // clip(lp[..., 0:1])   (:16,20)
        }
        for (int l = tid; l < wrow - 1; l += kEmitConsumers) {
            float e = 0.f;
            if (l < L) {
                const float z = (l == tid) ? zl : __ldg(rowp + __ldg(lab + l));
                const float lp = (z - M) - logS;             // log_softmax          (:123 / :14)
                e = ctc ? fmaxf(lp + add, kClip)             // clip(lp + log_voiced) (:131-132)
                        : fmaxf(lp, kClip);                  // clip(lp)             (:18)
            }
            Erow[1 + l] = e;
        }
        if (tid == 0) Erow[0] = blank;
    }
}

// LOGP mode: the caller already holds log-probs (the run_viterbi_core boundary); gather only.
__global__ void gather_logp_kernel(const EmitParams p) {
    const int64_t row = blockIdx.x;
    const int64_t grow = row + p.row0;
    int a = 0, b = p.m.n_utt;
    while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (p.m.t_off[mid] <= grow) a = mid; else b = mid;
    }
    const int l0 = p.m.l_off[a];
    const int L = p.m.l_off[a + 1] - l0;
    const int wrow = p.m.e_row[a];
    float* Erow = p.E + p.m.e_off[a] + (grow - p.m.t_off[a]) * (int64_t)wrow;
    const float* rowp = p.logits + row * p.ld;
    for (int l = threadIdx.x; l < wrow - 1; l += blockDim.x)
        Erow[1 + l] = (l < L) ? rowp[p.m.labels[l0 + l]] : 0.f;
    if (threadIdx.x == 0) Erow[0] = p.sil[row * p.ld_sil];
}

size_t emit_smem_bytes(int stages) {
    return (size_t)stages * kEmitStageBytes + 2 * stages * sizeof(uint64_t) + 2 * (kEmitConsumers / 32) * sizeof(float2);
}

template <int MODE, int STAGES, int CTAS>
static cudaError_t launch_emit_variant(const EmitParams& p, int sm_count, cudaStream_t stream) {
    const size_t smem = emit_smem_bytes(STAGES);
    static bool attr_done[64] = {};               // per-function, per-device opt-in: set once, not per launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(emit_kernel<MODE, STAGES, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    int grid = CTAS * sm_count;
    if (grid > p.n_rows) grid = p.n_rows;
    emit_kernel<MODE, STAGES, CTAS><<<grid, kEmitThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_emit(const EmitParams& p, int sm_count, cudaStream_t stream) {
    if (p.n_rows <= 0) return cudaSuccess;
    if (p.m.mode == 2) {
        gather_logp_kernel<<<p.n_rows, 128, 0, stream>>>(p);
        return cudaGetLastError();
    }
    // 3 CTAs/SM with 4-stage rings. Measured A/B on one box (profiles/k2_tuning_r1.md): with the SM clock
    // power-capped to ~1.7 GHz (the tensor-heavy K1 runs right before), 2 CTAs x 6 stages fell to 0.90 of the
    // HBM peak (consumer latency-bound), 3 CTAs x 4 stages holds 0.98; the slower shape is no longer built.
    return p.m.mode == 0 ? launch_emit_variant<0, 4, 3>(p, sm_count, stream) : launch_emit_variant<1, 4, 3>(p, sm_count, stream);
}

}  // namespace la
