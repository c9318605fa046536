"""Generates tests/golden/*.npz by RUNNING THE REFERENCE (only possible in the dev container,
where /root/reference is mounted):

    python tests/golden/make_golden.py

decode_cases.npz   -- inputs + outputs of the reference's perform_viterbi_ctc / perform_viterbi
                      (utils/alignment.py:13-71,121-188), the emission values its torch chain
                      produced at the label columns, and the dp score / 2-bit step codes of
                      its numba run_viterbi_core (:73-119).
core_cases.npz     -- run_viterbi_core on hand-made emission tables (exact ties, repeats,
                      floor stress): full dp (fp64) and bt (int64) tables.
mae_cases.npz      -- get_mae (:190-199) known answers.
logmel_hf.npz      -- log-mel of seeded waveforms from transformers' WhisperFeatureExtractor
                      (cross-check only; the reference's own dependency is absent).
The fixtures are committed; the GPU box never runs this script.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shim  # noqa: E402

ref = ref_shim.load()


def ref_emissions(pred: torch.Tensor, ctc: bool):
    """The reference's torch chain, line for line in meaning (utils/alignment.py:123-134 / :14-20)."""
    if ctc:
        lp = F.log_softmax(pred[:, :, 1:-1], dim=2)
        s = F.sigmoid(pred[:, :, -1:])
        emit = torch.clip(lp + torch.log(1.0 - s), min=-1000)
        blank = torch.clip(torch.log(s), min=-1000)
    else:
        lp = F.log_softmax(pred, dim=2)
        blank = torch.clip(lp[:, :, 0:1], min=-1000)
        emit = torch.clip(lp, min=-1000)[:, :, 1:]
    return emit, blank


def ref_tables(emit_i, blank_i, lab):
    T, S = emit_i.shape[0], 2 * len(lab) + 1
    dp = np.full((T, S), -10000000.0)
    bt = np.zeros((T, S), dtype=np.int64)
    e = emit_i.numpy()
    b = blank_i.numpy()
    dp[0][0] = b[0][0]
    dp[0][1] = e[0][lab[0] - 1]
    return ref.run_viterbi_core(dp, bt, e, b, lab)


def run_case(name, pred, labels, ctc, store):
    pred_t = torch.from_numpy(pred)
    fn = ref.perform_viterbi_ctc if ctc else ref.perform_viterbi
    lab_arg = torch.from_numpy(labels) if isinstance(labels, np.ndarray) else labels
    out = fn(pred_t, lab_arg)
    emit, blank = ref_emissions(pred_t, ctc)
    B = pred.shape[0]
    lens, flat, gathered, scores, codes = [], [], [], [], []
    for i in range(B):
        lab = np.array([int(x) for x in labels[i] if int(x) != -100], dtype=np.int64)
        lens.append(len(lab))
        flat.extend(out[i])
        gathered.append(emit[i].numpy()[:, lab - 1].reshape(-1))
        dp, bt = ref_tables(emit[i], blank[i], lab)
        end = dp.shape[1] - 1 if dp[-1][-1] > dp[-1][-2] else dp.shape[1] - 2
        scores.append(dp[-1][end])
        codes.append((np.arange(dp.shape[1])[None, :] - bt)[1:].astype(np.uint8).reshape(-1))
    if isinstance(labels, list):
        lmax = max(len(r) for r in labels)
        labels = np.array([list(r) + [-100] * (lmax - len(r)) for r in labels], dtype=np.int64)
    store[f"{name}/pred"] = pred
    store[f"{name}/labels"] = labels
    store[f"{name}/ctc"] = np.array(int(ctc))
    store[f"{name}/lens"] = np.array(lens, dtype=np.int64)
    store[f"{name}/onoff"] = np.array(flat, dtype=np.float64).reshape(-1, 2)
    store[f"{name}/emit_at_labels"] = np.concatenate(gathered).astype(np.float32)
    store[f"{name}/blank"] = blank.numpy().reshape(B, -1).astype(np.float32)
    store[f"{name}/score"] = np.array(scores, dtype=np.float64)
    store[f"{name}/codes"] = np.concatenate(codes)
    return out


def rand_labels(rng, B, lmin, lmax, vmax, p_repeat=0.15):
    rows = []
    for _ in range(B):
        L = int(rng.integers(lmin, lmax + 1))
        r = []
        for j in range(L):
            if j and rng.random() < p_repeat:
                r.append(r[-1])
            else:
                r.append(int(rng.integers(2, vmax + 1)))
        rows.append(r)
    m = max(len(r) for r in rows)
    return np.array([r + [-100] * (m - len(r)) for r in rows], dtype=np.int64)


def planted(rng, B, T, V, labels, ctc, scale=2.0, boost=6.0):
    """Random logits with the true label column boosted along a random monotone segmentation."""
    pred = (scale * rng.standard_normal((B, T, V))).astype(np.float32)
    sil_col = V - 1 if ctc else 0
    for i in range(B):
        lab = [int(x) for x in labels[i] if x != -100]
        cuts = np.sort(rng.choice(np.arange(1, T), size=min(2 * len(lab), T - 1), replace=False))
        seg = np.searchsorted(cuts, np.arange(T), side="right")        # 0..2L
        for t in range(T):
            k = min(seg[t], 2 * len(lab))
            if k % 2 == 1:
                pred[i, t, lab[k // 2]] += boost
                pred[i, t, sil_col] -= 3.0 if ctc else 0.0
            else:
                pred[i, t, sil_col] += boost if not ctc else 3.0
    return pred


def main():
    rng = np.random.default_rng(114514)
    store = {}
    # --- known answers recorded in SURVEY.md section 8(c) ---------------------------------
    V = 12
    z = np.zeros((1, 8, V), np.float32)
    run_case("tie_345", z, np.array([[3, 4, 5]]), True, store)
    run_case("tie_333", z, np.array([[3, 3, 3]]), True, store)
    run_case("min_T3", np.zeros((1, 3, V), np.float32), np.array([[3, 4, 5]]), True, store)
    run_case("min_T5_rep", np.zeros((1, 5, V), np.float32), np.array([[3, 3, 3]]), True, store)
    o = run_case("zero_b2", np.zeros((2, 6, V), np.float32), np.array([[3, 4, 5], [7, -100, -100]]), True, store)
    assert o == [[[0.06, 0.08], [0.08, 0.1], [0.1, 0.12]], [[0.1, 0.12]]], o
    sat = np.zeros((1, 4, V), np.float32); sat[..., -1] = 40.0
    o = run_case("sat_pos40", sat, np.array([[3, 4]]), True, store)
    assert o == [[[0.04, 0.06], [0.06, 0.08]]], o
    sat = np.zeros((1, 4, V), np.float32); sat[..., -1] = -120.0
    o = run_case("sat_neg120", sat, np.array([[3, 4]]), True, store)
    assert o == [[[0.0, 0.06], [0.06, 0.08]]], o
    o = run_case("t1_l1", np.zeros((1, 1, V), np.float32), np.array([[3]]), True, store)
    assert o == [[[0.0, 0.02]]], o
    run_case("tie_345_ce", z, np.array([[3, 4, 5]]), False, store)
    run_case("list_labels", np.zeros((2, 6, V), np.float32), [[3, 4, 5], [7]], True, store)
    # --- seeded random / planted, both flavours, ragged batches ---------------------------
    for ctc in (True, False):
        tag = "ctc" if ctc else "ce"
        for (B, T, Vv, lmin, lmax) in [(3, 40, 48, 1, 9), (2, 97, 64, 10, 20), (1, 250, 410, 12, 12),
                                      (2, 33, 40, 14, 16), (1, 140, 33, 33, 40)]:
            labels = rand_labels(rng, B, lmin, lmax, Vv - 2)
            pred = planted(rng, B, T, Vv, labels, ctc)
            run_case(f"planted_{tag}_B{B}_T{T}_V{Vv}", pred, labels, ctc, store)
            pred = (3.0 * rng.standard_normal((B, T, Vv))).astype(np.float32)
            run_case(f"random_{tag}_B{B}_T{T}_V{Vv}", pred, labels, ctc, store)
        # exact-tie stress: logits on a coarse grid
        labels = rand_labels(rng, 2, 5, 8, 30, p_repeat=0.3)
        pred = (rng.integers(-2, 3, size=(2, 30, 32)) * 0.5).astype(np.float32)
        run_case(f"grid_{tag}", pred, labels, ctc, store)
    # full-width vocabulary, tiny T
    labels = rand_labels(rng, 1, 3, 3, 402)
    pred = planted(rng, 1, 7, 21129, labels, True)
    run_case("fullV_ctc", pred, labels, True, store)
    pred = planted(rng, 1, 7, 21129, labels, False)
    run_case("fullV_ce", pred, labels, False, store)
    np.savez_compressed(os.path.join(HERE, "decode_cases.npz"), **store)

    # --- run_viterbi_core on hand-made emission tables ------------------------------------
    core = {}
    def core_case(name, T, ncols, lab, gen):
        lab = np.array(lab, dtype=np.int64)
        e = gen((T, ncols)).astype(np.float32)
        b = gen((T, 1)).astype(np.float32)
        S = 2 * len(lab) + 1
        dp = np.full((T, S), -10000000.0); bt = np.zeros((T, S), dtype=np.int64)
        dp[0][0] = b[0][0]; dp[0][1] = e[0][lab[0] - 1]
        dp, bt = ref.run_viterbi_core(dp, bt, e, b, lab)
        core[f"{name}/emit"] = e; core[f"{name}/blank"] = b; core[f"{name}/label"] = lab
        core[f"{name}/dp"] = dp; core[f"{name}/bt"] = bt
    core_case("ties_q8", 60, 20, [3, 4, 4, 7, 9, 9, 9, 2, 11], lambda s: rng.integers(-40, 1, size=s) / 8.0)
    core_case("all_equal", 24, 8, [1, 2, 3, 3, 5], lambda s: np.full(s, -1.5))
    core_case("all_clip", 400, 6, [2, 3, 4, 5], lambda s: np.full(s, -1000.0))
    core_case("random", 120, 50, list(rng.integers(1, 51, size=37)), lambda s: -5.0 * rng.random(s))
    core_case("wide", 90, 410, list(rng.integers(1, 403, size=70)), lambda s: -8.0 * rng.random(s))
    core_case("single", 9, 5, [4], lambda s: -rng.random(s))
    np.savez_compressed(os.path.join(HERE, "core_cases.npz"), **core)

    # --- get_mae ---------------------------------------------------------------------------
    v = ref.get_mae([[[0, .5], [.5, 1]]], [[[.02, .48], [.5, 1.02]]])
    assert v == 0.01500000000000001, v
    gts, prs, vals = [], [], []
    for _ in range(6):
        B = int(rng.integers(1, 4))
        gt = [[sorted(rng.random(2).tolist()) for _ in range(int(rng.integers(1, 7)))] for _ in range(B)]
        pr = [[[x + 0.02 * int(rng.integers(-5, 6)) for x in p] for p in u] for u in gt]
        gts.append(gt); prs.append(pr); vals.append(ref.get_mae(gt, pr))
    with open(os.path.join(HERE, "mae_cases.json"), "w") as f:
        json.dump({"gt": gts, "predict": prs, "mae": vals, "survey_known": v}, f)

    # --- log-mel cross-check with transformers --------------------------------------------
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor()
    lm = {}
    for name, n in [("n16000", 16000), ("n40123", 40123), ("n3333", 3333)]:
        t = np.arange(n) / 16000.0
        a = 0.1 * rng.standard_normal(n)
        for h in range(1, 6):
            a += (0.3 / h) * np.sin(2 * np.pi * 220 * h * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t))
        a[int(0.9 * n):] = 0.0
        a = a.astype(np.float32)
        # HF pads to 30 s; its per-sample max equals the global max for one sample. Pad with zeros
        # ourselves so the max-8 floor sees the same tensor: compare on the padded clip.
        padded = np.zeros(480000, np.float32); padded[:n] = a
        feat = fe._np_extract_fbank_features(padded[None], "cpu")[0]
        lm[f"{name}/audio"] = a
        lm[f"{name}/hf_logmel_padded30s"] = feat.astype(np.float32)[:, : n // 160 + 4]
    np.savez_compressed(os.path.join(HERE, "logmel_hf.npz"), **lm)
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
