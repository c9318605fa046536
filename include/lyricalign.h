/*
 * lyricalign.h -- C ABI of the B200-native alignment decode path (liblyricalign.so).
 *
 * This is the drop-in boundary for navi0105/LyricAlignment's decode path. The reference is
 * pure Python and has no FFI layer of its own; the interface each entry point replaces is a
 * Python call site (reference file:line cited per function). INTEGRATION.md shows the ctypes
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types. `stream` is a cudaStream_t
 *     passed as void* (NULL = legacy default stream).
 *   - Every function returns 0 on success or a negative la_status; nothing throws.
 *   - The big buffers (emission matrix, packed backpointers, K1 scratch) live in caller-provided
 *     workspaces whose sizes la_plan_workspace_bytes() / la_logmel_workspace_bytes() report; the
 *     device-pointer entry points (la_emit, la_viterbi, la_align, la_logmel*) enqueue on the caller's
 *     stream and allocate nothing.
 *   - Cached state (everything the library keeps between calls; la_shutdown() releases it all):
 *       * plan metadata: a few KB of device memory per plan, taken from a per-device pool of 64 KiB
 *         blocks by la_plan_create() and returned by la_plan_destroy();
 *       * K1's constant DFT basis table (520 KB per device, built on the first la_logmel* call);
 *       * the host-path context of la_align_host() (two staging buffers, a workspace, a pinned result
 *         buffer and two streams per device; grown on demand, calls on one device are serialised).
 *     There is no other process-wide mutable state; entry points are re-entrant across threads/streams.
 *   - Lifetime: la_plan_destroy() may be called while work enqueued with the plan is still running --
 *     the plan's metadata block is only recycled after an event recorded on the last stream the plan
 *     was used on has completed. The WORKSPACE and the in/out buffers are the caller's and must outlive
 *     the enqueued work as usual.
 *   - There is NO CPU fallback. Without a CUDA device every compute entry returns
 *     LA_ERR_CUDA.
 *
 * Data layout (one "batch" = n_utt independent utterances, ragged):
 *   logits   float32 [sum(T_u)][V] rows, row stride `ld` floats; utterance u owns rows
 *            [t_off[u], t_off[u+1]).  A padded [B][T][V] tensor is the special case T_u = T.
 *   labels   int32 [sum(L_u)], already resolved to ORIGINAL logit columns (label c reads
 *            column c: utils/alignment.py:86 `cur_log_prediction[j][cur_label[0] - 1]` on
 *            the `[:, :, 1:-1]` / `[:, :, 1:]` slice).
 *   outputs  first[sum(L_u)], last_plus1[sum(L_u)] int32 frame indices (onset = first*hop,
 *            offset = last_plus1*hop, utils/alignment.py:182-185), score[n_utt] fp64 =
 *            dp[T-1][end state], status[n_utt] (la_utt_status).
 */
#ifndef LYRICALIGN_H_
#define LYRICALIGN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum la_status {
    LA_OK = 0,
    LA_ERR_ARG = -1,      /* bad argument (null pointer, V too small, misaligned base, ...) */
    LA_ERR_CUDA = -2,     /* CUDA runtime error or no device; see la_last_error() */
    LA_ERR_LIMIT = -3,    /* utterance exceeds a compiled-in limit (L > LA_MAX_LABELS) */
    LA_ERR_ALLOC = -4
} la_status;

/* per-utterance status written by la_viterbi(); the Python mirror maps them to the
 * reference's exceptions */
typedef enum la_utt_status {
    LA_UTT_OK = 0,
    LA_UTT_EMPTY = 1,       /* no labels  -> reference raises IndexError (utils/alignment.py:152) */
    LA_UTT_INFEASIBLE = 2   /* a label state is not on the path -> ValueError (:183) */
} la_utt_status;

/* emission flavours */
#define LA_MODE_CTC 0       /* utils/alignment.py:123-134  (perform_viterbi_ctc) */
#define LA_MODE_CE 1        /* utils/alignment.py:14-20    (perform_viterbi, the "DTW" config) */
#define LA_MODE_LOGP 2      /* caller supplies log-probs: the run_viterbi_core boundary (:73-119) */

#define LA_MAX_LABELS 8191  /* longest label row one CTA can hold (32 warps x 32 lanes x 8 pairs) */

typedef struct la_plan la_plan;

const char* la_version(void);
const char* la_last_error(void);            /* thread-local, human readable */
/* Frees every cache listed under "Cached state" above on all devices (waits for the host-path streams
 * first). Live plans stay valid. Safe to call more than once; the caches are rebuilt on demand. */
void la_shutdown(void);
int la_device_count(void);

/* ---- plan: shapes + labels of one batch ------------------------------------------------
 * Replaces the per-utterance bookkeeping of utils/alignment.py:140-152 (label stripping,
 * dp/bt allocation). h_t_len[u] = frames, h_l_len[u] = labels of utterance u; h_labels is the
 * concatenation of the label rows (original logit columns, each in [0, V)).
 * For LA_MODE_LOGP, V is the number of columns of the caller's log-prob matrix and h_labels
 * holds column indices into it. */
int la_plan_create(la_plan** plan, int mode, int n_utt, int V, const int32_t* h_t_len,
                   const int32_t* h_l_len, const int32_t* h_labels, int device);
void la_plan_destroy(la_plan* plan);
size_t la_plan_workspace_bytes(const la_plan* plan);
int64_t la_plan_total_frames(const la_plan* plan);
int64_t la_plan_total_labels(const la_plan* plan);
/* kernels one la_align() enqueues for this plan: 1 (K2) + one K3 launch per non-empty L bucket */
int la_plan_num_launches(const la_plan* plan);
/* introspection for tests: where utterance u's emission rows / packed backpointers live in
 * the workspace. emit: float32 [T_u][row_floats], column 0 = blank/silence, column 1+l =
 * label l. bp: uint32 [word_rows][pairs_padded] (see la_plan_utt_bp_layout), one nibble per pair per frame =
 * code(blank state 2i) | code(label state 2i+1) << 1, code = k - backpointer in {0,1,2}. */
int la_plan_utt_layout(const la_plan* plan, int utt, int64_t* emit_off_bytes, int32_t* row_floats,
                       int64_t* bp_off_bytes, int32_t* pairs_padded);
/* introspection for tests, second half: the packed table is uint32 [word_rows = ceil(T_u/8)][pairs_padded]; nibble
 * t%8 of word row t/8 is frame t; pair i lives in column i + col_shift (1 for utterances of 33..639 pairs, where a
 * lane of the wavefront kernel owns the pairs 2j-1 and 2j; else 0). */
int la_plan_utt_bp_layout(const la_plan* plan, int utt, int32_t* word_rows, int32_t* col_shift);

/* ---- K2: fused log-softmax + label gather ----------------------------------------------
 * Replaces utils/alignment.py:123-134 (mode CTC) / :14-20 (mode CE): one streaming pass over
 * the [sum T][V] logits, never materialising the [T][V] log-softmax. For LA_MODE_LOGP,
 * d_logits is the caller's log-prob matrix, d_sil its [sum T] silence column (row stride
 * ld_sil floats) and the kernel only gathers. d_logits must be 16-byte aligned. */
int la_emit(const la_plan* plan, const float* d_logits, int64_t ld, const float* d_sil,
            int64_t ld_sil, void* d_workspace, void* stream);

/* ---- K3: Viterbi DP + packed backpointers + backtrace ----------------------------------
 * Replaces run_viterbi_core (utils/alignment.py:73-119), the end-state pick and backtrace
 * (:157-176) and the on/offset scan (:182-185). fp64 state, reference tie order, finite
 * -1e7 floor; indices are bit-exact given identical emissions. One launch per non-empty size
 * class of the plan: utterances of up to 63 / 255 / 639 state pairs (L + 1) run the lane-skewed
 * wavefront kernel on 1 / up to 4 / up to 10 warps, longer ones (L <= LA_MAX_LABELS) the
 * row-synchronous kernel; the packed backpointer table has the same layout either way. */
int la_viterbi(const la_plan* plan, void* d_workspace, int32_t* d_first, int32_t* d_last_plus1,
               double* d_score, int32_t* d_status, void* stream);

/* parity instrumentation: same as la_viterbi() on a 1-utterance plan, additionally dumping the
 * full fp64 DP table d_dp[T][2L+1] (what the reference keeps in `dp_matrix`) so tests can compare
 * every cell with run_viterbi_core's output. Not for production use. */
int la_viterbi_debug(const la_plan* plan, void* d_workspace, int32_t* d_first, int32_t* d_last_plus1,
                     double* d_score, int32_t* d_status, double* d_dp, void* stream);

/* ---- K2 + K3 in one call (device buffers) ----------------------------------------------
 * Replaces perform_viterbi_ctc / perform_viterbi (utils/alignment.py:121-188 / :13-71) up to
 * the final int -> seconds multiply, which stays in Python fp64. */
int la_align(const la_plan* plan, const float* d_logits, int64_t ld, void* d_workspace,
             int32_t* d_first, int32_t* d_last_plus1, double* d_score, int32_t* d_status,
             void* stream);

/* ---- N1: the head's Linear fused with K2 (device buffers) -------------------------------------
 * Replaces the producer AND the consumer of the [T][V] logits: `self.fc(self.activate(out))`
 * (module/align_model.py:32-33,38,107) and the emission math (utils/alignment.py:123-134 / :14-20). The caller
 * passes the Mish output d_X float32 [sum T][D] (row stride ldx; D a multiple of 32, <= 1024; 768 in the
 * reference), the Linear's weight d_W float32 [V][D] (row stride ldw) and bias d_bias [V]; the logits are never
 * materialised. la_head_pack_weights() converts the weight once per model into the tensor-core layout
 * (la_head_packed_weight_bytes(V, D) bytes of device memory, caller-owned). la_head_emit() fills the plan's
 * emission rows exactly as la_emit() does (so la_viterbi() follows); d_head_ws needs
 * la_head_workspace_bytes(plan, D) bytes. The normaliser is computed on the tensor cores from fp16 hi/lo slices
 * (|x|, |w| < 65504), the gathered label logits in plain fp32; emissions agree with the fp64 evaluation of
 * X W^T + b to 1e-4 (la_emit on materialised fp32 logits: 2e-5). */
size_t la_head_packed_weight_bytes(int V, int D);
int la_head_pack_weights(const float* d_W, int64_t ldw, int V, int D, void* d_packed, void* stream);
size_t la_head_workspace_bytes(const la_plan* plan, int D);
int la_head_emit(const la_plan* plan, const float* d_X, int64_t ldx, int D, const float* d_W, int64_t ldw,
                 const float* d_bias, const void* d_packed_w, void* d_head_workspace, void* d_workspace,
                 void* stream);
/* la_head_emit + la_viterbi */
int la_head_align(const la_plan* plan, const float* d_X, int64_t ldx, int D, const float* d_W, int64_t ldw,
                  const float* d_bias, const void* d_packed_w, void* d_head_workspace, void* d_workspace,
                  int32_t* d_first, int32_t* d_last_plus1, double* d_score, int32_t* d_status, void* stream);

/* ---- same, HOST buffers (the reference's actual call: logits already `.cpu()`ed,
 * inference_alignment.py:161-166). Streams the logits through a double-buffered device
 * staging area owned by a per-device library context (grown on first use and reused across
 * plans; `staging_bytes` each, 0 = default 64 MiB) overlapping H2D copies with K2; results are
 * copied back before returning. Calls on one device are serialised by the context's mutex.
 * h_logits should be pinned for full PCIe rate. */
int la_align_host(la_plan* plan, const float* h_logits, int64_t ld, int32_t* h_first,
                  int32_t* h_last_plus1, double* h_score, int32_t* h_status,
                  size_t staging_bytes);

/* ---- K1: log-mel front end --------------------------------------------------------------
 * Replaces whisper.audio.log_mel_spectrogram as called at module/align_model.py:84:
 * d_wave float32 [batch][n_samples] (row stride wave_stride floats; the reference zero-pads every
 * clip to the batch maximum, align_model.py:78-82) -> d_out float32 [batch][80][out_stride] with
 * the first n_samples/160 frames of every mel row written (columns beyond that are left
 * untouched, so the caller can pre-zero a 3000-frame encoder window, align_model.py:89). The
 * max-8 floor uses the GLOBAL maximum over the whole call, as whisper does. d_wave must be
 * 16-byte aligned; n_samples > 200 (reflect padding).
 * Workspace: la_logmel_workspace_bytes(n_clips, total_samples) bytes of device memory. */
size_t la_logmel_workspace_bytes(int n_clips, int64_t total_samples);
int la_logmel(const float* d_wave, int batch, int64_t n_samples, int64_t wave_stride, float* d_out,
              int64_t out_stride, void* d_workspace, void* stream);

/* Ragged form: n_clips INDEPENDENT calls of the above with batch 1 (the reference's default
 * --batch-size 1: every clip gets its own maximum) in one launch. Clip c reads
 * d_wave[h_wave_off[c] .. + h_n_samples[c]) and writes d_out[h_out_off[c] + m*h_out_stride[c] + f].
 * Offsets that are multiples of 4 floats take the TMA path. The h_* arrays are host memory. */
int la_logmel_ragged(const float* d_wave, int n_clips, const int64_t* h_wave_off,
                     const int32_t* h_n_samples, float* d_out, const int64_t* h_out_off,
                     const int32_t* h_out_stride, void* d_workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LYRICALIGN_H_ */
