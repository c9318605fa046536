"""CPU, dev container only: differential test of oracle/ against the reference module imported
unmodified from /root/reference (skipped where that tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

import oracle
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def _rand_case(rng, quantised):
    import torch
    B = int(rng.integers(1, 4))
    T = int(rng.integers(1, 60))
    V = int(rng.integers(6, 40))
    rows = []
    for _ in range(B):
        L = int(rng.integers(1, 12))
        r = []
        for j in range(L):
            r.append(r[-1] if j and rng.random() < 0.25 else int(rng.integers(1, V - 1)))
        rows.append(r)
    m = max(map(len, rows))
    labels = np.array([r + [-100] * (m - len(r)) for r in rows], dtype=np.int64)
    if quantised:
        pred = (rng.integers(-3, 4, size=(B, T, V)) * 0.25).astype(np.float32)
    else:
        pred = (2.5 * rng.standard_normal((B, T, V))).astype(np.float32)
    return torch.from_numpy(pred), torch.from_numpy(labels)


@pytest.mark.parametrize("quantised", [False, True])
@pytest.mark.parametrize("ctc", [True, False])
def test_differential_decode(ref, ctc, quantised):
    rng = np.random.default_rng(7 + 2 * ctc + quantised)
    rfn = ref.perform_viterbi_ctc if ctc else ref.perform_viterbi
    ofn = oracle.perform_viterbi_ctc if ctc else oracle.perform_viterbi
    n_ok = 0
    for _ in range(60):
        pred, labels = _rand_case(rng, quantised)
        try:
            want = rfn(pred, labels)
        except (ValueError, IndexError) as e:
            with pytest.raises(type(e)):
                ofn(pred.numpy(), labels.numpy())
            continue
        if quantised:
            # exact ties everywhere: libm-vs-Sleef ulps in the emissions may legitimately flip
            # them, so feed the oracle DP the reference's own emission values instead
            import torch, torch.nn.functional as F
            if ctc:
                lp = F.log_softmax(pred[:, :, 1:-1], dim=2)
                s = F.sigmoid(pred[:, :, -1:])
                emit = torch.clip(lp + torch.log(1.0 - s), min=-1000).numpy()
                blank = torch.clip(torch.log(s), min=-1000).numpy()
            else:
                lp = F.log_softmax(pred, dim=2)
                blank = torch.clip(lp[:, :, 0:1], min=-1000).numpy()
                emit = np.ascontiguousarray(torch.clip(lp, min=-1000)[:, :, 1:].numpy())
            got = []
            for i in range(pred.shape[0]):
                lab = [int(x) for x in labels[i] if x != -100]
                r = oracle.align_one(emit[i], blank[i], lab)
                assert r["status"] == 0
                got.append([[float(int(f)) * 0.02, float(int(l)) * 0.02] for f, l in zip(r["first"], r["last_plus1"])])
        else:
            got = ofn(pred.numpy(), labels.numpy())
        assert got == want
        n_ok += 1
    assert n_ok > 20


def test_differential_get_mae(ref):
    rng = np.random.default_rng(3)
    for _ in range(20):
        gt = [[sorted(rng.random(2).tolist()) for _ in range(int(rng.integers(1, 9)))] for _ in range(int(rng.integers(1, 5)))]
        pr = [[[x + 0.02 * int(rng.integers(-5, 6)) for x in p] for p in u] for u in gt]
        assert oracle.get_mae(gt, pr) == ref.get_mae(gt, pr)
