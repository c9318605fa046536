"""CPU: pins oracle/ against the fixtures the REFERENCE produced (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import oracle
import sys

from conftest import GOLDEN, split_onoff

GOLDEN_DIR = GOLDEN


def _labels_rows(labels):
    return [[int(x) for x in row] for row in labels]


def test_decode_golden_onoff_bit_exact(decode_cases):
    assert len(decode_cases) >= 30
    for name, c in decode_cases.items():
        fn = oracle.perform_viterbi_ctc if int(c["ctc"]) else oracle.perform_viterbi
        got = fn(c["pred"], c["labels"])
        want = split_onoff(c["onoff"], c["lens"])
        assert got == want, name          # Python floats, == is bit-exact


def test_decode_golden_emissions_and_scores(decode_cases):
    for name, c in decode_cases.items():
        ctc = bool(int(c["ctc"]))
        emit, blank = (oracle.emission_ctc if ctc else oracle.emission_ce)(c["pred"])
        emit64, blank64 = (oracle.emission_ctc_f64 if ctc else oracle.emission_ce_f64)(c["pred"])
        got, got64, scores, tol = [], [], [], []
        for i, row in enumerate(c["labels"]):
            lab = np.array([x for x in row if x != -100], dtype=np.int64)
            got.append(emit[i][:, lab - 1].reshape(-1))
            got64.append(emit64[i][:, lab - 1].reshape(-1))
            # fp32 `log(1 - sigmoid(z))` carries an absolute error of ~2^-24 (1 + e^z): the
            # reference's own chain is that far from exact arithmetic (hard part 4)
            zs = c["pred"][i][:, -1].astype(np.float64) if ctc else np.zeros(c["pred"].shape[1])
            tol.append(np.repeat(2e-5 + 2.0 ** -23 * (1.0 + np.exp(zs)), len(lab)))
            scores.append(oracle.align_one(emit[i], blank[i], lab)["score"])
        got, got64, tol = np.concatenate(got), np.concatenate(got64), np.concatenate(tol)
        sat = name.startswith("sat_")
        # fp32 restatement vs the reference's torch chain: libm-vs-Sleef ulps only
        np.testing.assert_allclose(got, c["emit_at_labels"], rtol=0, atol=5e-6, err_msg=name)
        np.testing.assert_allclose(blank.reshape(blank.shape[0], -1), c["blank"], rtol=0, atol=5e-6, err_msg=name)
        if not sat:   # fp64 anchor: the reference's own fp32 chain sits within 2e-5 of it
            assert np.all(np.abs(got64 - c["emit_at_labels"]) <= tol), name
        np.testing.assert_allclose(scores, c["score"], rtol=1e-6, err_msg=name)


def test_oracle_dp_on_reference_emissions_bit_exact(decode_cases):
    """Given the reference's own emission values (label columns only), the C oracle must
    reproduce its 2-bit step codes and fp64 score exactly."""
    for name, c in decode_cases.items():
        T = c["pred"].shape[1]
        pe, pc = 0, 0
        for i, row in enumerate(c["labels"]):
            lab = np.array([x for x in row if x != -100], dtype=np.int64)
            L, S = len(lab), 2 * len(lab) + 1
            e = c["emit_at_labels"][pe:pe + T * L].reshape(T, L); pe += T * L
            r = oracle.align_one(np.ascontiguousarray(e), c["blank"][i], np.arange(1, L + 1), want_tables=True)
            # repeats are encoded in the labels, so re-run with a label vector that keeps them
            uniq = {}
            rel = np.array([uniq.setdefault(int(v), len(uniq) + 1) for v in lab], dtype=np.int64)
            e2 = np.zeros((T, len(uniq)), np.float32)
            for j, v in enumerate(rel):
                e2[:, v - 1] = e[:, j]
            r = oracle.align_one(e2, c["blank"][i], rel, want_tables=True)
            codes = (np.arange(S)[None, :] - r["bt"])[1:].astype(np.uint8).reshape(-1)
            want = c["codes"][pc:pc + (T - 1) * S]; pc += (T - 1) * S
            assert np.array_equal(codes, want), name
            assert r["score"] == c["score"][i], name


def test_core_tables_bit_exact(core_cases):
    for name, c in core_cases.items():
        e, b, lab = c["emit"], c["blank"], c["label"]
        T, S = e.shape[0], 2 * len(lab) + 1
        dp = np.full((T, S), -10000000.0)
        bt = np.zeros((T, S), dtype=np.int64)
        dp[0][0] = b[0][0]
        dp[0][1] = e[0][lab[0] - 1]
        oracle.viterbi_core(dp, bt, e, b, lab)
        assert np.array_equal(dp, c["dp"]), name
        assert np.array_equal(bt, c["bt"]), name


def test_known_paths():
    """tie-breaking pushes every advance as late as possible."""
    z = np.zeros((8, 10), np.float32)
    b = np.zeros((8, 1), np.float32)
    assert oracle.align_one(z, b, [3, 4, 5])["path"].tolist() == [0, 0, 0, 0, 0, 1, 3, 5]
    assert oracle.align_one(z, b, [3, 3, 3])["path"].tolist() == [0, 0, 0, 1, 2, 3, 4, 5]
    assert oracle.align_one(z[:3], b[:3], [3, 4, 5])["path"].tolist() == [1, 3, 5]
    assert oracle.align_one(z[:5], b[:5], [3, 3, 3])["path"].tolist() == [1, 2, 3, 4, 5]


def test_error_behaviour():
    z = np.zeros((1, 4, 12), np.float32)
    with pytest.raises(IndexError):
        oracle.perform_viterbi_ctc(z, np.array([[-100, -100]]))
    with pytest.raises(ValueError):
        oracle.perform_viterbi_ctc(z, np.array([[3, 4, 5, 6, 7]]))      # T < L
    with pytest.raises(ValueError):
        oracle.perform_viterbi_ctc(z[:, :2], np.array([[3, 3]]))         # needs a blank between repeats
    assert oracle.perform_viterbi_ctc(z[:, :1], np.array([[3]])) == [[[0.0, 0.02]]]


def test_get_mae_golden():
    with open(os.path.join(GOLDEN, "mae_cases.json")) as f:
        g = json.load(f)
    assert oracle.get_mae([[[0, .5], [.5, 1]]], [[[.02, .48], [.5, 1.02]]]) == g["survey_known"] == 0.01500000000000001
    for gt, pr, v in zip(g["gt"], g["predict"], g["mae"]):
        assert oracle.get_mae(gt, pr) == v


def test_decode_frames_bankers_rounding():
    assert oracle.decode_frames(501) == 250 and oracle.decode_frames(503) == 252
    assert oracle.decode_frames(3000) == 1500
    from oracle.logmel import decode_frames_chunked
    assert decode_frames_chunked(30000) == 15000
    assert decode_frames_chunked(3001) == 1500 + 0      # round(0.5) == 0
    assert decode_frames_chunked(3003) == 1500 + 2


def test_mel_filterbank_matches_transformers():
    tf = pytest.importorskip("transformers.audio_utils")
    hf = tf.mel_filter_bank(201, 80, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T
    np.testing.assert_allclose(oracle.mel_filterbank(), hf, atol=1e-7)


def test_logmel_vs_hf_golden(logmel_cases):
    for name, c in logmel_cases.items():
        a = c["audio"]
        padded = np.zeros(480000, np.float32)
        padded[:len(a)] = a
        got = oracle.log_mel_spectrogram(padded)
        want = c["hf_logmel_padded30s"]
        np.testing.assert_allclose(got[:, :want.shape[1]], want, atol=2e-5, err_msg=name)


def test_logmel_fp32_formulation_close_to_fp64():
    from oracle.logmel import log_mel_spectrogram_torch_f32
    rng = np.random.default_rng(1)
    a = (0.1 * rng.standard_normal((2, 8000))).astype(np.float32)
    x = oracle.log_mel_spectrogram(a)
    y = log_mel_spectrogram_torch_f32(a).numpy()
    assert x.shape == (2, 80, 50)
    np.testing.assert_allclose(x, y, atol=1e-4 / 4)   # 1e-4 in log10 domain == 2.5e-5 after /4


def test_generated_mel_table_matches_oracle_filterbank():
    """lyricalignment_b200/csrc/la_mel_table.inc (compile-time literals of the K1 epilogue) must be
    the oracle's (== librosa's == whisper's) 80 x 201 filterbank, bit for bit."""
    import re
    from conftest import ROOT
    txt = open(os.path.join(ROOT, "lyricalignment_b200", "csrc", "la_mel_table.inc")).read()
    rows = re.findall(r"X\((\d+), (\d+), (\S+?)f, (\S+?)f, (\d+)\)", txt)
    assert len(rows) == 201
    W = oracle.mel_filterbank()
    rebuilt = np.zeros_like(W)
    prev = 0
    for k, lo, w0, w1, nf in rows:
        k, lo, nf = int(k), int(lo), int(nf)
        assert nf == (0 if k == 0 else lo - prev) and 0 <= nf <= 2
        rebuilt[lo, k] = np.float32(float.fromhex(w0))
        if lo + 1 < 80:
            rebuilt[lo + 1, k] = np.float32(float.fromhex(w1))
        else:
            assert float.fromhex(w1) == 0.0
        prev = lo
    assert np.array_equal(rebuilt, W)
    ms = int(re.search(r"#define LA_MEL_MS (\d+)", txt).group(1))
    split = int(re.search(r"#define LA_MEL_SPLIT (\d+)", txt).group(1))
    assert ms == int(rows[split][1])


@pytest.mark.parametrize("name", ["c1_30s", "c4_300s"])
def test_logmel_oracle_pinned_to_two_independent_implementations(name):
    """openai-whisper is absent, so the log-mel oracle is a restatement. Pin it on the benchmark shapes
    (30 s clip, 5-minute song) to the two implementations of the same algorithm this image has
    (tests/golden/make_logmel_golden.py): HF's numpy extractor (fp64 inside; must agree to fp32 storage
    precision) and HF's torch.stft fp32 extractor (whisper's own formulation; agrees to fp32-FFT accuracy)."""
    sys.path.insert(0, GOLDEN_DIR)
    from make_logmel_golden import signal
    g = np.load(os.path.join(GOLDEN_DIR, "logmel_hf_long.npz"))
    stride = int(g[f"{name}/stride"])
    got = oracle.log_mel_spectrogram(signal(name))[:, ::stride]
    e_np = 4.0 * np.abs(got - g[f"{name}/hf_numpy"])                # log10 units
    e_t = 4.0 * np.abs(got - g[f"{name}/hf_torch_f32"])
    assert e_np.max() <= 1e-6, float(e_np.max())
    assert e_t.max() <= 2e-4 and np.quantile(e_t, 0.9999) <= 1e-5, (float(e_t.max()), float(np.quantile(e_t, 0.9999)))


def test_k1_basis_table_slices_sum_to_the_windowed_dft_basis():
    """Host logic of K1 (no GPU): the fp16 operand slices la_logmel.cu builds for the tensor cores.
    B1 must lie on the 64-grid with |B1| <= 2^14 (that is what makes the Acc0 chain exact in fp32), and
    B1 + B2 + B3 must reproduce 2^14 w[n] cos|(-sin)(2 pi k n / 400) to 2^-30 of full scale."""
    import ctypes
    from lyricalignment_b200 import _lib
    lib = _lib.load()
    lib.la_debug_logmel_basis.restype = ctypes.c_size_t
    lib.la_debug_logmel_basis.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    n = lib.la_debug_logmel_basis(None, 0)
    assert n == 26 * 3 * 3328
    buf = np.zeros(n, np.uint16)
    lib.la_debug_logmel_basis(buf.ctypes.data, n)
    blocks = buf.view(np.float16).astype(np.float64).reshape(2, 13, 3, 2, 26, 8, 8)   # [pass][ks][slice][ki][ni][bin%8][kk%8]
    B = blocks.transpose(0, 2, 1, 3, 6, 4, 5).reshape(2, 3, 208, 208)                 # [pass][slice][n = 16ks + 8ki + kk][bin]
    nn = np.arange(208)
    w = np.where((nn >= 1) & (nn <= 200), 0.5 - 0.5 * np.cos(2 * np.pi * nn / 400.0), 0.0)
    ph = np.outer(nn, np.arange(208)) % 400
    C = w[:, None] * np.cos(2 * np.pi * ph / 400.0)
    S = -w[:, None] * np.sin(2 * np.pi * ph / 400.0)
    C[200] *= 0.5
    S[200] = 0.0
    C[:, 201:] = 0.0
    S[:, 201:] = 0.0
    for pas, want in ((0, C), (1, S)):
        b1, b2, b3 = B[pas]
        assert np.all(b1 % 64 == 0) and np.abs(b1).max() <= 16384
        assert np.abs(b2).max() <= 32.0 and np.abs(b3).max() <= 2.0 ** -6
        assert np.abs(b1 + b2 + b3 - 16384.0 * want).max() <= 2.0 ** -16      # 2^-30 of full scale
