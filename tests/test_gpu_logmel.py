"""GPU: K1 (tcgen05 log-mel) against the fp64 oracle restatement.

Tolerance. BASELINE.json:north_star asks for 1e-4 absolute in the log10 domain (== 2.5e-5 after the final /4) and
that is what every case asserts, on EVERY cell, the 5-minute song included -- no percentile allowance. Measured on
B200 (scripts/diag_logmel*.py, profiles/logmel_precision_r2.txt): worst cell 8.0e-6 on the 5-minute song, 1.0e-5 on
noise; the reference's own fp32 torch.stft path sits 2.7e-5 .. 8.2e-5 from the same oracle. (Round 1's 3xTF32 chain
with the tensor core's truncating accumulate was at 3.3e-4; round 2 slices the operands on a fixed grid so that the
leading chain is exact in fp32 -- see la_logmel.cu.) p99.9 is pinned an order of magnitude below the bar as well."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

from lyricalignment_b200 import audio as LA    # noqa: E402

TOL = 1e-4 / 4.0          # worst cell, output units == 1e-4 in the log10 domain
TOL_P999 = 1e-5 / 4.0     # 99.9th percentile


def _close(got, want):
    e = np.abs(got - want)
    assert e.max() <= TOL, float(e.max())
    if e.size >= 4000:
        assert np.quantile(e, 0.999) <= TOL_P999, float(np.quantile(e, 0.999))


def _signal(rng, n, kind):
    t = np.arange(n) / 16000.0
    if kind == "noise":
        a = 0.1 * rng.standard_normal(n)
    elif kind == "tone":
        a = 0.8 * np.sin(2 * np.pi * 440.0 * t)
    elif kind == "silence_tail":
        a = 0.1 * rng.standard_normal(n)
        a[int(0.6 * n):] = 0.0
    else:   # noise + 10 amplitude-modulated harmonics, last 10 % zeroed
        a = 0.1 * rng.standard_normal(n)
        for h in range(1, 11):
            a += (0.3 / h) * np.sin(2 * np.pi * 220 * h * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t))
        a[int(0.9 * n):] = 0.0
    return a.astype(np.float32)


@pytest.mark.parametrize("kind", ["noise", "tone", "silence_tail", "survey"])
@pytest.mark.parametrize("n", [16000, 40123, 3333, 480000, 20480 + 200, 641])
def test_single_clip_vs_fp64_oracle(kind, n):
    rng = np.random.default_rng(n % 97)
    a = _signal(rng, n, kind)
    got = LA.log_mel_spectrogram(a).cpu().numpy()
    want = oracle.log_mel_spectrogram(a)
    assert got.shape == want.shape == (80, n // 160)
    _close(got, want)


def test_batch_shares_global_max_and_padding():
    rng = np.random.default_rng(1)
    a = np.stack([_signal(rng, 48000, "survey"), 0.01 * _signal(rng, 48000, "noise"), np.zeros(48000, np.float32)])
    got = LA.log_mel_spectrogram(torch.from_numpy(a).cuda()).cpu().numpy()
    want = oracle.log_mel_spectrogram(a)                      # global max over the whole batch
    _close(got, want)
    assert np.all(got[2] == got[2, 0, 0])                     # digital silence sits on the floor
    got_p = LA.log_mel_spectrogram(a[0], padding=480).cpu().numpy()
    _close(got_p, oracle.log_mel_spectrogram(a[0], padding=480))


def test_ragged_launch_equals_independent_calls():
    rng = np.random.default_rng(2)
    lens = [16000 * 5 + 37, 16000 * 9, 3333, 16000 * 15 - 1, 801]
    offs, pos = [], 0
    for n in lens:
        offs.append(pos)
        pos += (n + 3) // 4 * 4
    wave = np.zeros(pos + 8, np.float32)
    clips = []
    for o, n in zip(offs, lens):
        c = _signal(rng, n, "survey") * rng.uniform(0.05, 1.0)
        wave[o:o + n] = c
        clips.append(c)
    out, ooff, frames = LA.log_mel_spectrogram_ragged(torch.from_numpy(wave).cuda(), offs, lens)
    out = out.cpu().numpy()
    for c, o, f in zip(clips, ooff, frames):
        got = out[o:o + 80 * f].reshape(80, f)
        _close(got, oracle.log_mel_spectrogram(c))


def test_writes_into_prezeroed_encoder_window():
    """module/align_model.py:89: pad_or_trim(mel, 3000) -- K1 can write straight into the padded window."""
    rng = np.random.default_rng(3)
    a = _signal(rng, 16000 * 7, "survey")
    mel = LA.log_mel_spectrogram(a)
    padded = LA.pad_or_trim(mel, LA.N_FRAMES)
    assert padded.shape == (80, 3000) and torch.all(padded[:, 700:] == 0)
    assert LA.decode_frames(mel.shape[-1]) == 350 and LA.decode_frames(501) == 250 and LA.decode_frames(503) == 252


def test_five_minute_song_and_chunked_framing():
    """BASELINE config 3 front end: 300 s = 30 000 mel frames = 10 encoder chunks (align_model.py:93-104);
    the max-8 floor is global over the whole song."""
    rng = np.random.default_rng(5)
    a = _signal(rng, 16000 * 300, "survey")
    got = LA.log_mel_spectrogram(a).cpu().numpy()
    assert got.shape == (80, 30000)
    # 2.4 M cells, every one of them within the north_star bound (measured worst: 8.0e-6)
    e = 4.0 * np.abs(got - oracle.log_mel_spectrogram(a))
    assert e.max() <= 1e-4 and np.quantile(e, 0.9999) <= 1e-5, (float(e.max()), float(np.quantile(e, 0.9999)))
    assert sum(LA.decode_frames(min(3000, 30000 - s)) for s in range(0, 30000, 3000)) == 15000


def test_unaligned_clip_offsets_take_the_plain_load_path():
    """Clip starts that are not multiples of 4 samples cannot use TMA; results must not change."""
    rng = np.random.default_rng(6)
    clips = [_signal(rng, n, "noise") for n in (20001, 33333, 16000)]
    offs, pos = [], 3
    for c in clips:
        offs.append(pos)
        pos += len(c) + 1          # odd gaps -> every alignment class
    wave = np.zeros(pos + 8, np.float32)
    for o, c in zip(offs, clips):
        wave[o:o + len(c)] = c
    out, ooff, frames = LA.log_mel_spectrogram_ragged(torch.from_numpy(wave).cuda(), offs, [len(c) for c in clips])
    out = out.cpu().numpy()
    for c, o, f in zip(clips, ooff, frames):
        _close(out[o:o + 80 * f].reshape(80, f), oracle.log_mel_spectrogram(c))
