// la_head.cu -- N1: the head's Linear(D -> V) fused with the online log-sum-exp and the label gather, so the
// [T][V] fp32 logits (84.5 KB per 20 ms frame at V = 21129) are never written or read.
//
// Replaces, in one go, the producer  `self.fc(self.activate(out))`   (module/align_model.py:32-33,38,107)
// and the consumer                    log_softmax / sigmoid / clip    (utils/alignment.py:123-134, :14-20)
// of that tensor: the caller hands over the Mish output X [sum T][D] (D = 768), the Linear's weight W [V][D] and
// bias b [V]; out come the same compact emission rows K2 writes (blank + L label columns per frame), ready for K3.
//
// Three kernels:
//   head_pack_kernel    fp32 -> fp16 hi/lo slices in the UMMA canonical K-major layout, one contiguous block per
//                       (tile, k-step), so the GEMM's operands arrive by plain 1-D bulk copies. W is packed once
//                       per model (la_head_pack_weights), X once per call (1 ms for a 2 000-clip batch).
//   head_lse_kernel     tcgen05 GEMM, M = 128 rows x N = 256 vocabulary columns per tile, K = 16 per MMA, both
//                       operands from shared memory. x = hi + lo with hi, lo fp16 (22 significant bits), and
//                       z = hi*hi + (hi*lo + lo*hi): the main term and the cross terms accumulate in SEPARATE TMEM
//                       accumulators (the tensor core's fp32 accumulate truncates; see la_logmel.cu), summed in
//                       round-to-nearest in the epilogue. A persistent CTA owns a row tile and sweeps all 83 column
//                       tiles, so the epilogue keeps a per-row online (max, sum exp) pair in registers -- the
//                       logits go TMEM -> registers -> two floats per row. All CTAs sweep the column tiles in the
//                       same order, so the 65 MB of packed weights are served from L2. A call with fewer row
//                       tiles than SMs (one clip = 4 tiles) splits the column sweep over several CTAs per row tile;
//                       the gather kernel merges the partial (max, sum) pairs.
//   head_gather_kernel  the <= L + 1 columns a frame actually needs (its utterance's labels + the silence / class-0
//                       column), as exact fp32 dot products on the CUDA cores (63 GFLOP for the whole batch), then
//                       the reference's emission formulas in its operation order -- the same epilogue as K2.
// Accuracy: the normaliser carries the split-fp16 GEMM's error (~1e-6 relative per logit, plus the truncation of a
// 48-deep accumulate chain, ~3e-6 of |z|); the gathered label logits are plain fp32. Stated tolerance of the
// emissions against the fp64 oracle: 1e-4 (tests/test_gpu_head.py); K2's, on materialised logits, is 2e-5.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>

#include "la_common.cuh"

namespace la {

constexpr int kHM = 128;                    // rows per tile
constexpr int kHN = 256;                    // vocabulary columns per tile
constexpr int kHK = 16;                     // K per MMA (fp16)
constexpr int kHStages = 6;
constexpr int kHABlk = kHM * kHK * 2;       // 4096 B: one fp16 slice of an A k-step  [ki 2][mi 16][8 rows][8 fp16]
constexpr int kHBBlk = kHN * kHK * 2;       // 8192 B: one fp16 slice of a  B k-step  [ki 2][ni 32][8 rows][8 fp16]
constexpr int kHStageBytes = 2 * kHABlk + 2 * kHBBlk;    // 24576: [A_hi A_lo B_hi B_lo]
constexpr int kHThreads = 384;              // warp 0 producer, warp 1 MMA issuer, warps 4..11 epilogue
constexpr float kHLog2e = 1.4426950408889634f;

__device__ __forceinline__ float h_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void h_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void h_tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void h_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void h_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// both operands from shared memory; one elected lane of a converged warp issues
__device__ __forceinline__ void h_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void h_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void h_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void h_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// UMMA shared-memory descriptor, K-major, no swizzle (same fields as la_logmel.cu's)
__device__ __forceinline__ uint64_t h_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// D = f32, A = B = f16, K-major, N >> 3 @ [17,23), M >> 4 @ [24,29)
constexpr uint32_t kHIdesc = (1u << 4) | ((uint32_t)(kHN >> 3) << 17) | ((uint32_t)(kHM >> 4) << 24);

// ---------------------------------------------------------------------------------------------
// packing: src fp32 [rows][D] (row stride ld) -> [tile][k-step][hi | lo] canonical blocks of R rows x 16
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void head_pack_kernel(const float* __restrict__ src, int64_t ld, int rows, int D, unsigned char* __restrict__ dst,
                                 int tiles) {
    const int ksteps = D / kHK;
    const int64_t total = (int64_t)tiles * ksteps * R;
    constexpr int kBlk = R * kHK * 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i % R);
        const int64_t tk = i / R;
        const int ks = (int)(tk % ksteps);
        const int64_t tile = tk / ksteps;
        const int64_t row = tile * R + r;
        float v[16];
        if (row < rows) {
            const float4* s4 = reinterpret_cast<const float4*>(src + row * ld + ks * kHK);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = __ldg(s4 + q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = 0.f;
        }
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const __half2 h = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn(v[2 * q] - f.x, v[2 * q + 1] - f.y);
            hi[q] = *reinterpret_cast<const uint32_t*>(&h);
            lo[q] = *reinterpret_cast<const uint32_t*>(&l);
        }
        unsigned char* blk = dst + (size_t)tk * (2 * kBlk);
        const int off = (r >> 3) * 128 + (r & 7) * 16;                   // + ki * (R / 8 * 128)
        *reinterpret_cast<uint4*>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(blk + (R / 8) * 128 + off) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(blk + kBlk + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(blk + kBlk + (R / 8) * 128 + off) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// GEMM + online log-sum-exp
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kHThreads, 1) head_lse_kernel(const HeadParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stages = smem;
    float* bias_s = reinterpret_cast<float*>(smem + kHStages * kHStageBytes);     // [2][256], masked columns = -inf
    float2* comb = reinterpret_cast<float2*>(bias_s + 2 * kHN);                    // [128] second half's (max, sum)
    uint64_t* full = reinterpret_cast<uint64_t*>(comb + kHM);                      // [kHStages]
    uint64_t* empty = full + kHStages;                                             // [kHStages]
    uint64_t* tmem_full = empty + kHStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kHStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 8);
        mbar_fence_init();
    }
    if (warp == 1) h_tmem_alloc(tmem_slot, 512);
    h_fence_before();
    __syncthreads();
    h_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== producer ==========================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int w = blockIdx.x; w < p.m_tiles * p.n_splits; w += gridDim.x) {
                const int mt = w / p.n_splits, sp = w % p.n_splits;
                const int nt1 = min(p.n_tiles, (sp + 1) * p.n_per_split);
                for (int nt = sp * p.n_per_split; nt < nt1; ++nt)
                    for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
                        const int s = it % kHStages;
                        mbar_wait(&empty[s], ((it / kHStages) & 1) ^ 1);
                        mbar_arrive_expect_tx(&full[s], kHStageBytes);
                        unsigned char* dst = stages + s * kHStageBytes;
                        bulk_g2s(dst, p.xp + ((size_t)mt * p.ksteps + ks) * (2 * kHABlk), 2 * kHABlk, &full[s]);
                        bulk_g2s(dst + 2 * kHABlk, p.wp + ((size_t)nt * p.ksteps + ks) * (2 * kHBBlk), 2 * kHBBlk, &full[s]);
                    }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (whole warp, one elected lane issues) ====
        const uint32_t acc_main = tmem_base, acc_cross = tmem_base + kHN;
        const uint32_t s0 = smem_u32(stages);
        uint32_t it = 0, tc = 0;
        for (int w = blockIdx.x; w < p.m_tiles * p.n_splits; w += gridDim.x) {
            const int sp = w % p.n_splits;
            const int nt1 = min(p.n_tiles, (sp + 1) * p.n_per_split);
            for (int nt = sp * p.n_per_split; nt < nt1; ++nt, ++tc) {
                mbar_wait(tmem_empty, (tc & 1) ^ 1);               // the epilogue has read the previous tile out of TMEM
                h_fence_after();
                for (int ks = 0; ks < p.ksteps; ++ks, ++it) {
                    const int s = it % kHStages;
                    mbar_wait(&full[s], (it / kHStages) & 1);
                    h_fence_after();
                    const uint32_t base = s0 + s * kHStageBytes;
                    const uint64_t a_hi = h_desc(base, kHM / 8 * 128, 128), a_lo = h_desc(base + kHABlk, kHM / 8 * 128, 128);
                    const uint64_t b_hi = h_desc(base + 2 * kHABlk, kHN / 8 * 128, 128);
                    const uint64_t b_lo = h_desc(base + 2 * kHABlk + kHBBlk, kHN / 8 * 128, 128);
                    const uint32_t acc = ks > 0 ? 1u : 0u;
                    h_mma_ss(acc_cross, a_hi, b_lo, kHIdesc, acc);
                    h_mma_ss(acc_main, a_hi, b_hi, kHIdesc, acc);
                    h_mma_ss(acc_cross, a_lo, b_hi, kHIdesc, 1u);
                    h_commit(&empty[s]);
                }
                h_commit(tmem_full);
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue: online (max, sum exp) per row ==============
        // thread = row = TMEM lane; warps 4..7 take columns [0,128) of the tile, warps 8..11 [128,256)
        const int q = warp & 3, half = warp >= 8 ? 1 : 0;
        const int row = q * 32 + lane;
        const int e = tid - 128;                                    // 0..255: which bias column this thread stages
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
        uint32_t tc = 0;
        for (int w = blockIdx.x; w < p.m_tiles * p.n_splits; w += gridDim.x) {
            const int mt = w / p.n_splits, sp = w % p.n_splits;
            const int nt1 = min(p.n_tiles, (sp + 1) * p.n_per_split);
            float m = -INFINITY, s = 0.f;
            for (int nt = sp * p.n_per_split; nt < nt1; ++nt, ++tc) {
                {
                    const int col = nt * kHN + e;
                    bias_s[(tc & 1) * kHN + e] = (col >= p.col_lo && col <= p.col_hi) ? __ldg(p.bias + col) : -INFINITY;
                }
                named_bar_sync(1, 256);                             // this tile's bias row is staged (double-buffered by tile parity)
                mbar_wait(tmem_full, tc & 1);
                h_fence_after();
                const float4* bs = reinterpret_cast<const float4*>(bias_s + (tc & 1) * kHN + half * 128);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    uint32_t a[16], b[16];
                    h_ld16(lane_base + 16 * c, a);
                    h_ld16(lane_base + kHN + 16 * c, b);
                    h_ld_wait();
                    if (c == 7) {                                   // this warp's share of TMEM is in registers
                        h_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tmem_empty);
                    }
                    float x[16];
                    float cmax = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bb = bs[4 * c + j];            // same address in every lane: broadcast
                        x[4 * j] = (__uint_as_float(a[4 * j]) + __uint_as_float(b[4 * j])) + bb.x;
                        x[4 * j + 1] = (__uint_as_float(a[4 * j + 1]) + __uint_as_float(b[4 * j + 1])) + bb.y;
                        x[4 * j + 2] = (__uint_as_float(a[4 * j + 2]) + __uint_as_float(b[4 * j + 2])) + bb.z;
                        x[4 * j + 3] = (__uint_as_float(a[4 * j + 3]) + __uint_as_float(b[4 * j + 3])) + bb.w;
                        cmax = fmaxf(cmax, fmaxf(fmaxf(x[4 * j], x[4 * j + 1]), fmaxf(x[4 * j + 2], x[4 * j + 3])));
                    }
                    if (cmax > -INFINITY) {
                        const float mn = fmaxf(m, cmax);
                        float acc = s * h_ex2((m - mn) * kHLog2e);
                        const float nm = -mn * kHLog2e;
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc += h_ex2(fmaf(x[j], kHLog2e, nm));
                        s = acc;
                        m = mn;
                    }
                }
            }
            // the two column halves of a row meet here
            if (half) comb[row] = make_float2(m, s);
            named_bar_sync(1, 256);
            if (!half) {
                const float2 o = comb[row];
                const float M = fmaxf(m, o.x);
                const float S = (m > -INFINITY ? s * expf(m - M) : 0.f) + (o.x > -INFINITY ? o.y * expf(o.x - M) : 0.f);
                const int64_t grow = (int64_t)mt * kHM + row;
                if (grow < p.rows) p.lse[(int64_t)sp * p.rows + grow] = make_float2(M, S);
            }
            named_bar_sync(1, 256);                                 // comb is free for the next row tile
        }
    }

    h_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        h_tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// exact fp32 logits of the columns a frame needs + the reference's emission formulas
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) head_gather_kernel(const HeadGatherParams g) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = g.D >> 5;                                        // floats per lane (D is a multiple of 32, <= 1024)
    for (int rr = 0; rr < 4; ++rr) {
        const int64_t row = (int64_t)blockIdx.x * 32 + warp * 4 + rr;
        if (row >= g.rows) return;
        int a = 0, b = g.m.n_utt;                                   // last u with t_off[u] <= row
        while (b - a > 1) {
            const int mid = (a + b) >> 1;
            if (__ldg(&g.m.t_off[mid]) <= row) a = mid; else b = mid;
        }
        const int l0 = __ldg(&g.m.l_off[a]);
        const int L = __ldg(&g.m.l_off[a + 1]) - l0;
        const int wrow = __ldg(&g.m.e_row[a]);
        float* Erow = g.E + __ldg(&g.m.e_off[a]) + (row - __ldg(&g.m.t_off[a])) * (int64_t)wrow;
        const int32_t* lab = g.m.labels + l0;
        float x[32];
        const float* xr = g.X + row * g.ldx;
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = (i < nk) ? __ldg(xr + lane + 32 * i) : 0.f;
        auto logit = [&](int col) -> float {                        // every lane returns the full dot product + bias
            const float* w = g.W + (int64_t)col * g.ldw;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < nk) acc = fmaf(x[i], __ldg(w + lane + 32 * i), acc);
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            return acc + __ldg(g.bias + col);
        };
        float M = -INFINITY, S = 0.f;                                // the column splits of the row's normaliser meet here
        for (int sp = 0; sp < g.n_splits; ++sp) M = fmaxf(M, g.lse[(int64_t)sp * g.rows + row].x);
        for (int sp = 0; sp < g.n_splits; ++sp) {
            const float2 ms = g.lse[(int64_t)sp * g.rows + row];
            if (ms.x > -INFINITY) S += ms.y * expf(ms.x - M);
        }
        const float logS = logf(S);
        const float zsil = logit(MODE == 0 ? g.m.V - 1 : 0);
        float add = 0.f, blank;
        if (MODE == 0) {
            const float sg = 1.0f / (1.0f + expf(-zsil));           // F.sigmoid            (utils/alignment.py:125)
            add = logf(1.0f - sg);                                  // log(1 - s)           (:126,129)
            blank = fmaxf(logf(sg), kClip);                         // clip(log s, -1000)   (:128,134)
        } else {
            blank = fmaxf((zsil - M) - logS, kClip);                // clip(lp[..., 0:1])   (:16,20)
        }
        for (int c0 = 0; c0 < wrow - 1; c0 += 32) {
            float z = 0.f;
            const int n = min(32, L - c0);
            for (int j = 0; j < n; ++j) {
                const float v = logit(__ldg(lab + c0 + j));
                if (lane == j) z = v;
            }
            const int l = c0 + lane;
            if (l < wrow - 1) {
                float eo = 0.f;
                if (l < L) {
                    const float lp = (z - M) - logS;                // log_softmax          (:123 / :14)
                    eo = MODE == 0 ? fmaxf(lp + add, kClip)         // clip(lp + log_voiced) (:131-132)
                                   : fmaxf(lp, kClip);              // clip(lp)             (:18)
                }
                Erow[1 + l] = eo;
            }
        }
        if (lane == 0) Erow[0] = blank;
    }
}

size_t head_lse_smem_bytes() { return (size_t)kHStages * kHStageBytes + 2 * kHN * 4 + kHM * 8 + (2 * kHStages + 2) * 8 + 16; }
int head_tile_rows() { return kHM; }
int head_tile_cols() { return kHN; }
size_t head_packed_bytes(int64_t rows, int D, bool weights) {
    const int R = weights ? kHN : kHM;
    const int64_t tiles = (rows + R - 1) / R;
    return (size_t)tiles * (D / kHK) * 2 * R * kHK * 2;
}

cudaError_t launch_head_pack(const float* src, int64_t ld, int64_t rows, int D, void* dst, bool weights, cudaStream_t stream) {
    const int R = weights ? kHN : kHM;
    const int64_t tiles = (rows + R - 1) / R;
    if (tiles == 0) return cudaSuccess;
    const int64_t total = tiles * (D / kHK) * R;
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
    if (weights) head_pack_kernel<kHN><<<grid, 256, 0, stream>>>(src, ld, (int)rows, D, static_cast<unsigned char*>(dst), (int)tiles);
    else head_pack_kernel<kHM><<<grid, 256, 0, stream>>>(src, ld, (int)rows, D, static_cast<unsigned char*>(dst), (int)tiles);
    return cudaGetLastError();
}

cudaError_t launch_head_lse(const HeadParams& p, int sm_count, cudaStream_t stream) {
    if (p.m_tiles <= 0) return cudaSuccess;
    const size_t smem = head_lse_smem_bytes();
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(head_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    head_lse_kernel<<<std::min(p.m_tiles * p.n_splits, sm_count), kHThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_head_gather(const HeadGatherParams& g, cudaStream_t stream) {
    if (g.rows <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((g.rows + 31) / 32);
    if (g.m.mode == 0) head_gather_kernel<0><<<grid, 256, 0, stream>>>(g);
    else head_gather_kernel<1><<<grid, 256, 0, stream>>>(g);
    return cudaGetLastError();
}

}  // namespace la
