"""GPU: waveform -> K1 -> stock encoder + head -> K2/K3, with the reference's framing arithmetic."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe():
    from lyricalignment_b200.pipeline import AlignPipeline
    return AlignPipeline("tiny", vocab=410, device="cuda")


def _wave(rng, n):
    t = np.arange(n) / 16000.0
    return (0.1 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 220 * t)).astype(np.float32)


@pytest.mark.parametrize("n_samples,want_T", [(16000 * 5, 250), (501 * 160, 250), (503 * 160, 252), (480000, 1500),
                                              (480000 + 160 * 3, 1502), (16000 * 47, 2350)])
def test_decode_frames_follow_reference_rounding(pipe, n_samples, want_T):
    """module/align_model.py:87-105: T = round-half-even(F/2) per 3000-frame chunk."""
    rng = np.random.default_rng(n_samples % 1000)
    logits = pipe.frame_manual_forward([_wave(rng, n_samples)])
    assert logits.shape == (1, want_T, 410) and logits.is_cuda
    assert want_T == oracle.logmel.decode_frames_chunked(n_samples // 160)


def test_pipeline_logits_decode_like_the_oracle(pipe):
    import lyricalignment_b200 as la
    rng = np.random.default_rng(7)
    audios = [_wave(rng, 16000 * 6), _wave(rng, 16000 * 4 + 77)]       # ragged: second clip zero-padded
    logits = pipe.frame_manual_forward(audios)
    assert logits.shape[:2] == (2, 300)
    labels = torch.tensor([[5, 9, 9, 33, 120, 7], [40, 41, 42, -100, -100, -100]])
    got = la.perform_viterbi_ctc(logits, labels)                        # stays on the device
    want = oracle.perform_viterbi_ctc(logits.cpu().numpy(), labels.numpy())
    assert got == want
    assert la.perform_viterbi_ctc(logits.cpu(), labels) == want         # the reference's .cpu() call shape
