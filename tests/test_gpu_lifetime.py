"""GPU: lifetime rules of the C ABI (include/lyricalign.h, "Cached state" / "Lifetime").

* la_plan_destroy() while the plan's kernels are still queued on a NON-BLOCKING stream, immediately followed by
  la_plan_create() of a different plan: the pooled metadata block must not be handed over before the first plan's
  work has finished (round-1 ADVICE: it was, and the legacy-stream cudaMemcpy of the new metadata does not order
  against a cudaStreamNonBlocking stream).
* la_shutdown() frees every cache; the next call rebuilds them."""
import ctypes

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

from lyricalignment_b200 import _lib, alignment as A       # noqa: E402
from lyricalignment_b200 import audio as LA                # noqa: E402


def _case(seed, B, T, V, L):
    rng = np.random.default_rng(seed)
    pred = (2.0 * rng.standard_normal((B, T, V))).astype(np.float32)
    labels = [rng.integers(2, min(V - 2, 403), size=L).astype(np.int64) for _ in range(B)]
    return pred, labels


def test_plan_destroyed_and_recreated_while_its_kernels_are_queued():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    V = 5000
    side = torch.cuda.Stream(device=dev)                              # torch side streams are cudaStreamNonBlocking
    cases = [_case(s, 6, 700, V, 20 + s) for s in range(6)]
    wants = [oracle.perform_viterbi_ctc(p, [l.tolist() for l in lab]) for p, lab in cases]
    preds = [torch.from_numpy(p).to(dev) for p, _ in cases]
    torch.cuda.synchronize()
    outs = []
    spin = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    with torch.cuda.stream(side):
        for _ in range(20):
            spin.mul_(1.0001)                                         # keep the stream busy: the launches below only queue
        for (p, lab), z in zip(cases, preds):
            l_len, cols = A._resolve_columns(lab, V - 2)
            plan = A.AlignPlan(A.MODE_CTC, V, np.full(len(lab), p.shape[1], np.int32), l_len, cols, 0)
            ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
            first = torch.empty(plan.total_labels, dtype=torch.int32, device=dev)
            last = torch.empty_like(first)
            score = torch.empty(plan.n_utt, dtype=torch.float64, device=dev)
            status = torch.empty(plan.n_utt, dtype=torch.int32, device=dev)
            _lib.check(lib.la_align(plan.handle, z.data_ptr(), V, ws.data_ptr(), first.data_ptr(), last.data_ptr(),
                                    score.data_ptr(), status.data_ptr(), side.cuda_stream), "la_align")
            plan.close()                                              # destroyed with its work still queued
            outs.append((first, last, status, l_len, ws))
    side.synchronize()
    for (first, last, status, l_len, _), want in zip(outs, wants):
        res = A.AlignResult(first.cpu().numpy(), last.cpu().numpy(), np.zeros(len(l_len)), status.cpu().numpy(), l_len)
        assert A.onoff_seconds(res) == want


def test_shutdown_releases_the_caches_and_they_come_back():
    lib = _lib.load()
    pred, labels = _case(11, 2, 90, 3000, 7)
    want = oracle.perform_viterbi_ctc(pred, [l.tolist() for l in labels])
    wav = (0.1 * np.random.default_rng(0).standard_normal(16000)).astype(np.float32)
    mel0 = LA.log_mel_spectrogram(wav).cpu().numpy()
    assert A.perform_viterbi_ctc(torch.from_numpy(pred), labels) == want          # host path: builds its context
    torch.cuda.synchronize()
    lib.la_shutdown()
    lib.la_shutdown()                                                              # idempotent
    assert A.perform_viterbi_ctc(torch.from_numpy(pred), labels) == want          # context rebuilt on demand
    assert A.perform_viterbi_ctc(torch.from_numpy(pred).cuda(), labels) == want
    assert np.array_equal(LA.log_mel_spectrogram(wav).cpu().numpy(), mel0)         # basis table rebuilt, same bits
