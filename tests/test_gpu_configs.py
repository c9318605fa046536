"""GPU: the remaining BASELINE.json configs as parity / property tests.

config 2 (MIR-1k shape, on_offset MAE through the CE decoder -- the north star's "DTW" path),
config 1 at full batch shape (property checks: every clip feasible, on/offsets monotone, a random
subset bit-exact against the oracle), and the sharded path on one GPU (shard -> align -> stitch ==
unsharded)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

import lyricalignment_b200 as la                              # noqa: E402
from lyricalignment_b200 import alignment as A, sharded, synth   # noqa: E402


def _planted_truth(batch, seed):
    """Ground-truth on/offsets of synth.planted_logits' segmentation (same RNG stream)."""
    rng = np.random.default_rng(seed + 1)
    gt = []
    for T, lab in zip(batch.t_len, batch.labels):
        T, L = int(T), len(lab)
        ncut = min(2 * L, T - 1)
        cuts = np.sort(rng.choice(np.arange(1, T), size=ncut, replace=False))
        seg = np.minimum(np.searchsorted(cuts, np.arange(T), side="right"), 2 * L)
        g = []
        for l in range(L):
            idx = np.nonzero(seg == 2 * l + 1)[0]
            g.append([float(idx[0]) * 0.02, float(idx[-1] + 1) * 0.02])
        gt.append(g)
    return gt


def test_mir1k_shape_ce_decoder_mae():
    """17 songs, T = 1400..5400, L = 48..171 (the real MIR-1k annotated shapes), CE
    flavour; small V keeps the logits small. MAE against the planted ground truth must equal the
    oracle's MAE exactly (same alignments) and be small (the planted path is recovered)."""
    rng = np.random.default_rng(17)
    t_len = rng.integers(1400, 5401, size=17).astype(np.int32)
    labels = []
    for t in t_len:
        L = int(rng.integers(48, 172))
        ids = rng.integers(2, 403, size=L)
        labels.append(ids.astype(np.int64))
    batch = synth.ClipBatch(t_len * 0.02, t_len.astype(np.int64) * 320, t_len, labels)
    V = 410
    z = synth.planted_logits(batch, V, ctc=False, device="cuda", seed=5)
    res = la.align_clips(z, t_len, labels, mode=A.MODE_CE)
    got = la.onoff_seconds(res)
    zc = z.cpu().numpy()
    want, r0 = [], 0
    for t, lab in zip(t_len, labels):
        want += oracle.perform_viterbi(zc[None, r0:r0 + t], [lab.tolist()])
        r0 += int(t)
    assert got == want
    gt = _planted_truth(batch, 5)
    mae = la.get_mae(gt, got)
    assert mae == oracle.get_mae(gt, want)
    assert mae < 0.05, mae


def test_opencpop_batch_properties_and_sampled_parity():
    """BASELINE config 1 shape: 2 000 clips of 5-15 s (V reduced to 410 so the test fits any GPU; the
    full-width stream is exercised by bench.py and test_full_vocab_clip_vs_oracle_end_to_end)."""
    batch = synth.opencpop_shaped(2000, seed=99)
    V = 410
    z = synth.planted_logits(batch, V, ctc=True, device="cuda", seed=99)
    res = la.align_clips(z, batch.t_len, batch.labels)
    assert np.all(res.status == 0)
    p = 0
    for u, L in enumerate(res.l_len):
        f, l = res.first[p:p + L], res.last_plus1[p:p + L]
        assert np.all(l > f) and np.all(f[1:] >= l[:-1]) and f[0] >= 0 and l[-1] <= batch.t_len[u]
        p += L
    assert np.all(np.isfinite(res.score)) and np.all(res.score < 0)
    # bit-exact on a random subset
    offs = np.concatenate([[0], np.cumsum(batch.t_len)])
    loffs = np.concatenate([[0], np.cumsum(res.l_len)])
    zc = None
    for u in np.random.default_rng(0).choice(2000, size=25, replace=False):
        x = z[offs[u]:offs[u + 1]].cpu().numpy()
        want = oracle.perform_viterbi_ctc(x[None], [batch.labels[u].tolist()])[0]
        got = [[float(a) * 0.02, float(b) * 0.02] for a, b in
               zip(res.first[loffs[u]:loffs[u + 1]], res.last_plus1[loffs[u]:loffs[u + 1]])]
        assert got == want, u
    # sharded (2 "ranks" on one GPU) == unsharded, in dataset order
    parts = []
    for r in range(2):
        lo, hi = sharded.shard_bounds(2000, 2, r)
        parts.append(la.align_clips(z[offs[lo]:offs[hi]], batch.t_len[lo:hi], batch.labels[lo:hi]))
    assert np.array_equal(np.concatenate([q.first for q in parts]), res.first)
    assert np.array_equal(np.concatenate([q.last_plus1 for q in parts]), res.last_plus1)
    assert np.array_equal(np.concatenate([q.score for q in parts]), res.score)
