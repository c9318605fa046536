"""oracle/logmel.py -- TEST INFRASTRUCTURE ONLY (CPU checker, never the product path).

Restatement of the reference's log-mel front end. The reference calls the THIRD-PARTY
``whisper.audio.log_mel_spectrogram`` (module/align_model.py:9,84; openai-whisper, unpinned in
requirements.txt:6, not vendored, not installed here), so this follows that package's
published algorithm:

    window = hann(400, periodic); stft = torch.stft(audio, 400, 160, window, center=True,
    pad_mode="reflect"); power = |stft[..., :-1]|**2; mel = filters(80x201, librosa Slaney) @ power;
    x = log10(clamp(mel, 1e-10)); x = max(x, x.max() - 8.0)  [GLOBAL max over the call tensor];
    out = (x + 4) / 4

and the reference's own framing arithmetic (module/align_model.py:87-105: frames -> decode
frames with Python's half-to-even ``round``). **Parity unpinned** against the reference repo
(it has no fixture for this); cross-checked against ``transformers``' WhisperFeatureExtractor
in tests/golden/make_golden.py -> tests/golden/logmel_hf.npz.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_MELS = 80
N_FRAMES = 3000


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_mels: int = N_MELS, n_fft: int = N_FFT, sr: int = SAMPLE_RATE) -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels=80) (Slaney scale, Slaney area norm),
    which is what whisper ships as assets/mel_filters.npz. float32 [n_mels, n_fft//2+1]."""
    n_bins = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    mel_f = _mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_bins), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]          # in-place on float32: product formed in fp64, rounded once
    return w


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def power_spectrogram_f64(audio: np.ndarray) -> np.ndarray:
    """[..., N] -> [..., 201, N//160] power, fp64, torch.stft(center=True, reflect) framing with
    the last frame dropped."""
    a = np.asarray(audio, dtype=np.float64)
    lead = a.shape[:-1]
    a = a.reshape(-1, a.shape[-1])
    n = a.shape[-1]
    if n <= N_FFT // 2:
        raise ValueError("reflect padding needs more than n_fft//2 samples")
    pad = np.pad(a, ((0, 0), (N_FFT // 2, N_FFT // 2)), mode="reflect")
    nf = n // HOP_LENGTH
    idx = (np.arange(nf) * HOP_LENGTH)[:, None] + np.arange(N_FFT)[None, :]
    frames = pad[:, idx] * hann_periodic()[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)
    power = (spec.real ** 2 + spec.imag ** 2).transpose(0, 2, 1)
    return power.reshape(*lead, N_FFT // 2 + 1, nf)


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = N_MELS, padding: int = 0) -> np.ndarray:
    """fp64 restatement; returns float64 [..., 80, N//160]. Global max over the whole call."""
    a = np.asarray(audio, dtype=np.float64)
    if padding > 0:
        a = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(0, padding)])
    power = power_spectrogram_f64(a)
    mel = np.matmul(mel_filterbank(n_mels).astype(np.float64), power)
    x = np.log10(np.maximum(mel, 1e-10))
    x = np.maximum(x, x.max() - 8.0)
    return (x + 4.0) / 4.0


def log_mel_spectrogram_torch_f32(audio, n_mels: int = N_MELS, padding: int = 0):
    """The reference's actual CPU formulation (torch.stft in fp32); used as the CPU baseline."""
    import torch
    import torch.nn.functional as F
    a = audio if torch.is_tensor(audio) else torch.from_numpy(np.asarray(audio))
    a = a.to(torch.float32)
    if padding > 0:
        a = F.pad(a, (0, padding))
    stft = torch.stft(a, N_FFT, HOP_LENGTH, window=torch.hann_window(N_FFT), return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    mel = torch.from_numpy(mel_filterbank(n_mels)) @ mag
    x = torch.clamp(mel, min=1e-10).log10()
    x = torch.maximum(x, x.max() - 8.0)
    return (x + 4.0) / 4.0


def decode_frames(n_mel_frames: int) -> int:
    """module/align_model.py:88,98: ``int(round(F / 2.0))`` -- Python rounds half to even."""
    return int(round(n_mel_frames / 2.0))


def decode_frames_chunked(n_mel_frames: int) -> int:
    """module/align_model.py:87-104: <= 3000 frames -> one window; longer -> independent
    3000-frame chunks each contributing round(chunk_len / 2) decode frames."""
    if n_mel_frames <= N_FRAMES:
        return decode_frames(n_mel_frames)
    return sum(decode_frames(min(N_FRAMES, n_mel_frames - s)) for s in range(0, n_mel_frames, N_FRAMES))
