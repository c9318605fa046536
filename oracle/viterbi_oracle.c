/*
 * oracle/viterbi_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU checker, never the product path).
 *
 * Plain-C restatement of the reference forced-alignment DP, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg to check the CUDA path.
 * Nothing under lyricalignment_b200/ may include, link or call this file.
 *
 * Reference followed (navi0105/LyricAlignment):
 *   utils/alignment.py:73-119   run_viterbi_core   -> la_oracle_viterbi_core()
 *   utils/alignment.py:141-152  dp/bt allocation and row-0 preset (-1e7 floor, fp64/int64)
 *   utils/alignment.py:157-176  strict end-state pick + backtrace
 *   utils/alignment.py:182-185  first/last occurrence of every odd (label) state
 * Pinned against the reference itself: tests/golden/*.npz were produced by importing
 * /root/reference/utils/alignment.py (tests/golden/make_golden.py) and
 * tests/test_oracle.py replays them through this file.
 *
 * Arithmetic contract: dp is IEEE fp64, one add per cell,
 * emissions are fp32 promoted exactly; comparisons are the reference's own
 * (`>` strict for stay-vs-previous, `>=` for the skip), evaluated in its branch order;
 * every state of every frame is computed (no pruning); unreachable cells start at the
 * finite floor -1e7, not -inf.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/liboracle.so oracle/viterbi_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LA_ORACLE_FLOOR (-10000000.0)

/* numpy/numba index semantics of `cur_log_prediction[j][cur_label[k//2] - 1]`
 * (utils/alignment.py:86): a negative index wraps once; anything else out of
 * range is an error (numba would read out of bounds -- we refuse instead). */
static int64_t resolve_col(int64_t label, int64_t ncols) {
    int64_t c = label - 1;
    if (c < 0) c += ncols;
    if (c < 0 || c >= ncols) return -1;
    return c;
}

/* utils/alignment.py:73-119. dp/bt are [T][S] row-major, row 0 preset by the caller.
 * logp is [T][*] with row stride ld_logp (floats), sil is [T][*] with row stride ld_sil. */
int la_oracle_viterbi_core(double* dp, int64_t* bt, const float* logp, int64_t ld_logp,
                           int64_t ncols, const float* sil, int64_t ld_sil,
                           const int64_t* label, int64_t L, int64_t T) {
    const int64_t S = 2 * L + 1;
    int64_t* col = (int64_t*)malloc(sizeof(int64_t) * (size_t)(L > 0 ? L : 1));
    if (!col) return -2;
    for (int64_t i = 0; i < L; ++i) {
        col[i] = resolve_col(label[i], ncols);
        if (col[i] < 0) { free(col); return 3; }
    }
    for (int64_t j = 1; j < T; ++j) {
        const double* p = dp + (j - 1) * S;
        double* q = dp + j * S;
        int64_t* b = bt + j * S;
        const float* lp = logp + j * ld_logp;
        const double blank = (double)sil[j * ld_sil];
        for (int64_t k = 0; k < S; ++k) {
            if (k == 0) {                                   /* :78-82 */
                b[k] = k;
                q[k] = p[k] + blank;
            } else if (k == 1) {                            /* :84-90 */
                const double e = (double)lp[col[0]];
                if (p[k] > p[k - 1]) { b[k] = k;     q[k] = p[k] + e; }
                else                 { b[k] = k - 1; q[k] = p[k - 1] + e; }
            } else if (k % 2 == 0) {                        /* :92-101 */
                if (p[k] > p[k - 1]) { b[k] = k;     q[k] = p[k] + blank; }
                else                 { b[k] = k - 1; q[k] = p[k - 1] + blank; }
            } else {                                        /* :103-117 */
                const double e = (double)lp[col[k / 2]];
                if (p[k - 2] >= p[k - 1] && p[k - 2] >= p[k] && label[k / 2] != label[k / 2 - 1]) {
                    b[k] = k - 2; q[k] = p[k - 2] + e;
                } else if (p[k] > p[k - 1]) {
                    b[k] = k;     q[k] = p[k] + e;
                } else {
                    b[k] = k - 1; q[k] = p[k - 1] + e;
                }
            }
        }
    }
    free(col);
    return 0;
}

/* utils/alignment.py:141-187 for ONE utterance, given its emission matrices.
 * Outputs (any may be NULL): path[T], first[L], last_plus1[L], score (dp[T-1][end]),
 * dp_out[T*S], bt_out[T*S].
 * Returns 0 ok, 1 empty label row (reference: IndexError, :152), 2 infeasible -- some
 * label state never appears on the backtraced path (reference: ValueError from
 * list.index, :183), 3 label column out of range, -2 out of memory. */
int la_oracle_align(const float* logp, int64_t ld_logp, int64_t ncols, const float* sil,
                    int64_t ld_sil, const int64_t* label, int64_t L, int64_t T,
                    int32_t* path, int32_t* first, int32_t* last_plus1, double* score,
                    double* dp_out, int64_t* bt_out) {
    if (L <= 0) return 1;
    if (T <= 0) return 2;
    const int64_t S = 2 * L + 1;
    const int64_t c0 = resolve_col(label[0], ncols);
    if (c0 < 0) return 3;
    double* dp = dp_out ? dp_out : (double*)malloc(sizeof(double) * (size_t)(T * S));
    int64_t* bt = bt_out ? bt_out : (int64_t*)malloc(sizeof(int64_t) * (size_t)(T * S));
    int32_t* pth = path ? path : (int32_t*)malloc(sizeof(int32_t) * (size_t)T);
    int rc = 0;
    if (!dp || !bt || !pth) { rc = -2; goto done; }
    for (int64_t i = 0; i < T * S; ++i) { dp[i] = LA_ORACLE_FLOOR; bt[i] = 0; }   /* :144,146 */
    dp[0] = (double)sil[0];                                                          /* :151 */
    dp[1] = (double)logp[c0];                                                        /* :152 */
    rc = la_oracle_viterbi_core(dp, bt, logp, ld_logp, ncols, sil, ld_sil, label, L, T);
    if (rc) goto done;
    {
        const double* last = dp + (T - 1) * S;
        int64_t k = (last[S - 1] > last[S - 2]) ? S - 1 : S - 2;                     /* :157 */
        if (score) *score = last[k];
        pth[T - 1] = (int32_t)k;
        int64_t cur = bt[(T - 1) * S + k];                                           /* :162,170 */
        for (int64_t j = T - 2; j >= 0; --j) {                                       /* :164-166 */
            pth[j] = (int32_t)cur;
            cur = bt[j * S + cur];
        }
        for (int64_t l = 0; l < L; ++l) {                                            /* :182-185 */
            const int32_t want = (int32_t)(2 * l + 1);
            int64_t f = -1, g = -1;
            for (int64_t j = 0; j < T; ++j) if (pth[j] == want) { f = j; break; }
            if (f < 0) { rc = 2; goto done; }
            for (int64_t j = T - 1; j >= 0; --j) if (pth[j] == want) { g = j; break; }
            if (first) first[l] = (int32_t)f;
            if (last_plus1) last_plus1[l] = (int32_t)(g + 1);
        }
    }
done:
    if (!dp_out) free(dp);
    if (!bt_out) free(bt);
    if (!path) free(pth);
    return rc;
}
