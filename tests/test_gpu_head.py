"""GPU: N1, the head's Linear fused with the emission math (la_head.cu) against the fp64 oracle.

Oracle: z = X W^T + b in fp64 (numpy), then the reference's emission math in fp64 (oracle.emission_*_f64) and the
C restatement of its DP. Stated tolerance of the emissions: 1e-4 (+ the fp32 sigmoid term K2's test also carries):
the normaliser comes from a split-fp16 tensor-core GEMM with a truncating accumulate, the gathered label logits
are plain fp32. The DP on the kernel's OWN emissions stays bit-exact, and on inputs with a clear winner the indices
equal those of perform_viterbi*(fc(hidden)) on materialised logits."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

import lyricalignment_b200 as la                                   # noqa: E402
from lyricalignment_b200 import alignment as A                     # noqa: E402
from lyricalignment_b200._lib import MODE_CE, MODE_CTC             # noqa: E402
from lyricalignment_b200.head import FusedHead                     # noqa: E402
from test_gpu_decode import _oracle_on_emissions, emission_tolerance   # noqa: E402


def _problem(seed, t_len, l_len, V, D, scale=1.0):
    rng = np.random.default_rng(seed)
    X = (scale * rng.standard_normal((int(np.sum(t_len)), D))).astype(np.float32)
    W = (rng.standard_normal((V, D)) / np.sqrt(D)).astype(np.float32)
    b = (0.5 * rng.standard_normal(V)).astype(np.float32)
    labels = [rng.integers(2, min(V - 2, 403), size=L).astype(np.int64) for L in l_len]
    return X, W, b, labels


def _check(X, W, b, labels, t_len, mode, tag):
    head = FusedHead(torch.from_numpy(W), torch.from_numpy(b))
    job = head.align_clips_async(torch.from_numpy(X).cuda(), t_len, labels, mode)
    res = job.result()
    z64 = X.astype(np.float64) @ W.astype(np.float64).T + b.astype(np.float64)
    ctc = mode == MODE_CTC
    r0, p = 0, 0
    worst = 0.0
    for u, lab in enumerate(labels):
        T, L = int(t_len[u]), len(lab)
        emis = A.unpack_emissions(job.plan, job.ws, u)
        zz = z64[r0:r0 + T]
        e64, b64 = (oracle.emission_ctc_f64 if ctc else oracle.emission_ce_f64)(zz)
        want = np.concatenate([b64, e64[:, np.asarray(lab) - 1]], axis=1)
        tol = 1e-4 - 2e-5 + emission_tolerance(zz[:, -1], ctc)[:, None]
        err = np.abs(emis - want)
        clipped = want <= -999.0
        assert np.all((err <= tol) | clipped), (tag, u, float(err[~clipped].max()))
        worst = max(worst, float(err[~clipped].max()))
        o = _oracle_on_emissions(emis, lab)                      # DP bit-exact on the kernel's own emissions
        assert int(res.status[u]) == o["status"], (tag, u)
        if o["status"] == 0:
            assert np.array_equal(res.first[p:p + L], o["first"]), (tag, u)
            assert np.array_equal(res.last_plus1[p:p + L], o["last_plus1"]), (tag, u)
        assert res.score[u] == o["score"], (tag, u)
        r0 += T
        p += L
    job.close()
    return worst


@pytest.mark.parametrize("mode", [MODE_CTC, MODE_CE])
def test_head_small_vocab_ragged(mode):
    t_len, l_len = [37, 5, 130, 64], [4, 1, 9, 12]
    X, W, b, labels = _problem(1, t_len, l_len, V=700, D=64)
    _check(X, W, b, labels, t_len, mode, f"small-{mode}")


def test_head_reference_shape_full_vocab():
    """D = 768, V = 21129 (train_multitask.py:657), a few clips: 83 column tiles, 2 row tiles."""
    t_len, l_len = [150, 97], [11, 7]
    X, W, b, labels = _problem(2, t_len, l_len, V=21129, D=768, scale=0.7)
    worst = _check(X, W, b, labels, t_len, MODE_CTC, "full")
    print("worst emission error vs fp64:", worst)


def test_head_many_row_tiles_and_large_logits():
    """More row tiles than SMs would need a big batch; 9 tiles exercise the persistent loop. Large |z| stresses the
    truncating accumulate (the error grows with |z|)."""
    t_len, l_len = [300, 411, 277, 128], [20, 33, 25, 3]
    X, W, b, labels = _problem(3, t_len, l_len, V=1500, D=256, scale=3.0)
    _check(X, W, b, labels, t_len, MODE_CTC, "tiles")


def test_head_equals_logits_path_on_planted_inputs():
    """perform_viterbi_ctc(hidden) == perform_viterbi_ctc(fc(hidden)) when the alignment has a clear winner:
    hidden states that point at their label's weight row along a planted segmentation."""
    rng = np.random.default_rng(4)
    V, D, T, L = 900, 128, 200, 9
    W = (rng.standard_normal((V, D)) / np.sqrt(D)).astype(np.float32)
    b = np.zeros(V, np.float32)
    lab = rng.choice(np.arange(2, 403), size=L, replace=False).astype(np.int64)
    seg = np.minimum((np.arange(T) * (2 * L + 1)) // T, 2 * L)
    X = (0.3 * rng.standard_normal((T, D))).astype(np.float32)
    for t in range(T):
        k = seg[t]
        if k % 2:
            X[t] += 6.0 * W[lab[k // 2]] / np.linalg.norm(W[lab[k // 2]])
            X[t] -= 3.0 * W[V - 1] / np.linalg.norm(W[V - 1])
        else:
            X[t] += 6.0 * W[V - 1] / np.linalg.norm(W[V - 1])
    head = FusedHead(torch.from_numpy(W), torch.from_numpy(b))
    hidden = torch.from_numpy(X).cuda().view(1, T, D)
    got = head.perform_viterbi_ctc(hidden, [lab.tolist()])
    logits = hidden @ head.weight.T + head.bias
    assert got == la.perform_viterbi_ctc(logits, [lab.tolist()])
    z64 = X.astype(np.float64) @ W.astype(np.float64).T
    assert got == oracle.perform_viterbi_ctc(z64.astype(np.float32)[None], [lab.tolist()])
    # host hidden states (3 KB per frame over PCIe instead of 84.5 KB of logits)
    assert head.perform_viterbi_ctc(torch.from_numpy(X).view(1, T, D), [lab.tolist()]) == got
