"""Drop-in for the ``whisper.audio`` names the reference imports (module/align_model.py:9:
``N_FRAMES, pad_or_trim, log_mel_spectrogram``), with the log-mel computed by the tcgen05 kernel
(K1) and returned ON THE GPU -- the reference computes it on the CPU and then ``.to(device)``
(align_model.py:84), which becomes a no-op."""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_FRAMES = N_SAMPLES // HOP_LENGTH      # 3000
N_MELS = 80


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """whisper.audio.pad_or_trim: zero-pad or cut `axis` to `length` (tensor or ndarray)."""
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad = [(0, 0)] * array.ndim
            pad[axis] = (0, length - array.shape[axis])
            array = F.pad(array, [p for sizes in pad[::-1] for p in sizes])
        return array
    if array.shape[axis] > length:
        array = array.take(indices=range(length), axis=axis)
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def _device_of(device) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device(device)
    if d.type != "cuda":
        raise _lib.LyricAlignError("lyricalignment_b200.log_mel_spectrogram runs on CUDA only (no CPU fallback)")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def log_mel_spectrogram(audio, n_mels: int = N_MELS, padding: int = 0, device=None) -> torch.Tensor:
    """whisper.audio.log_mel_spectrogram(audio, n_mels=80, padding=0, device=None):
    audio float32 [..., N] (ndarray or tensor, host or device) -> CUDA float32 [..., 80, N // 160].
    The max-8 floor uses the maximum over the WHOLE call tensor, as upstream does."""
    if n_mels != N_MELS:
        raise ValueError("only the 80-bin filterbank the reference uses is built")
    if not torch.cuda.is_available():
        raise _lib.LyricAlignError("lyricalignment_b200 needs a CUDA device (no CPU fallback)")
    lib = _lib.load()
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    dev = audio.device if (audio.is_cuda and device is None) else _device_of(device)
    x = audio.to(device=dev, dtype=torch.float32, non_blocking=True)
    if padding > 0:
        x = F.pad(x, (0, padding))
    lead = x.shape[:-1]
    n = x.shape[-1]
    x = x.reshape(-1, n).contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    B = x.shape[0]
    frames = n // HOP_LENGTH
    out = torch.empty((B, N_MELS, frames), dtype=torch.float32, device=dev)
    if B == 0 or frames == 0:
        return out.reshape(*lead, N_MELS, frames)
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.la_logmel_workspace_bytes(B, B * n)), dtype=torch.uint8, device=dev)
        _lib.check(lib.la_logmel(x.data_ptr(), B, n, x.stride(0), out.data_ptr(), frames, ws.data_ptr(),
                                 torch.cuda.current_stream(dev).cuda_stream), "la_logmel")
    return out.reshape(*lead, N_MELS, frames)


def log_mel_spectrogram_ragged(wave: torch.Tensor, offsets: Sequence[int], n_samples: Sequence[int],
                               out: torch.Tensor | None = None, out_offsets=None, out_strides=None):
    """Many independent clips (each one its own whisper call, i.e. its own maximum) in ONE launch.
    wave: float32 [total], CUDA or (pinned) host; clip c = wave[offsets[c] : offsets[c] + n_samples[c]].
    Returns (out, out_offsets, frames): out is a flat CUDA buffer, clip c's [80, frames[c]] matrix
    starts at out_offsets[c] with row stride out_strides[c] (= frames[c] by default)."""
    lib = _lib.load()
    assert wave.dtype == torch.float32 and wave.is_contiguous()
    if not wave.is_cuda:      # host waveforms (pin them): one async H2D, then the launch
        wave = wave.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    dev = wave.device
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    ns = np.ascontiguousarray(n_samples, dtype=np.int32)
    frames = (ns // HOP_LENGTH).astype(np.int32)
    if out_strides is None:
        out_strides = frames
    ostr = np.ascontiguousarray(out_strides, dtype=np.int32)
    if out_offsets is None:
        sizes = N_MELS * ostr.astype(np.int64)
        out_offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]) if len(sizes) else np.zeros(0, np.int64)
    ooff = np.ascontiguousarray(out_offsets, dtype=np.int64)
    if out is None:
        total = int((ooff[-1] + N_MELS * int(ostr[-1])) if len(ooff) else 0)
        out = torch.empty(total, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(int(lib.la_logmel_workspace_bytes(len(ns), int(ns.astype(np.int64).sum()))),
                         dtype=torch.uint8, device=dev)
        _lib.check(lib.la_logmel_ragged(wave.data_ptr(), len(ns), off.ctypes.data, ns.ctypes.data, out.data_ptr(),
                                        ooff.ctypes.data, ostr.ctypes.data, ws.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream), "la_logmel_ragged")
    return out, ooff, frames


def decode_frames(n_mel_frames: int) -> int:
    """module/align_model.py:88,98: int(round(F / 2.0)) with Python's half-to-even rounding."""
    return int(round(n_mel_frames / 2.0))
