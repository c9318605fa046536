"""Synthetic Opencpop-shaped workloads: there is no network for datasets or
checkpoints, so benchmarks and size-scaled tests use seeded clip shapes, labels and logits."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

V_HEAD = 21129           # len(bert-base-chinese tokenizer) + 1 (train_multitask.py:657)
N_CLASSES = 402          # pinyin classes in bert_base_chinese_pronunce_table.json
HOP_SECONDS = 0.02


@dataclass
class ClipBatch:
    durations: np.ndarray      # seconds, float64 [B]
    n_samples: np.ndarray      # int64 [B]
    t_len: np.ndarray          # int32 [B] decode frames
    labels: List[np.ndarray]   # class ids (1-based), per clip

    @property
    def audio_seconds(self) -> float:
        return float(self.durations.sum())


def decode_frames(n_samples: int) -> int:
    """module/align_model.py:88: mel frames F = N // 160, decode frames = int(round(F / 2.0))."""
    return int(round((n_samples // 160) / 2.0))


def opencpop_shaped(n_clips: int, seed: int = 114514, dmin: float = 5.0, dmax: float = 15.0) -> ClipBatch:
    rng = np.random.default_rng(seed)
    d = rng.uniform(dmin, dmax, size=n_clips)
    n = np.floor(16000 * d).astype(np.int64)
    t = np.array([decode_frames(int(x)) for x in n], dtype=np.int32)
    labels = []
    for i in range(n_clips):
        L = int(np.clip(round(2.4 * d[i] * rng.uniform(0.7, 1.3)), 1, max(1, t[i] // 2)))
        ids = rng.integers(2, N_CLASSES + 1, size=L)
        rep = rng.random(L) < 0.05
        for j in range(1, L):
            if rep[j]:
                ids[j] = ids[j - 1]
        labels.append(ids.astype(np.int64))
    return ClipBatch(d, n, t, labels)


def fresh_labels(batch: ClipBatch, seed: int) -> List[np.ndarray]:
    """New lyrics for the SAME clips (same durations / frame counts): what a recycled logits pool is aligned
    against in the 10^6-clip run. Vectorised: 2 000 clips in well under a millisecond."""
    rng = np.random.default_rng(seed)
    n = len(batch.t_len)
    L = np.clip(np.rint(2.4 * batch.durations * rng.uniform(0.7, 1.3, size=n)), 1, np.maximum(1, batch.t_len // 2)).astype(np.int64)
    ids = rng.integers(2, N_CLASSES + 1, size=int(L.sum()))
    rep = rng.random(ids.size) < 0.05
    starts = np.concatenate([[0], np.cumsum(L)[:-1]])
    rep[starts] = False
    idx = np.nonzero(rep)[0]
    ids[idx] = ids[idx - 1]                                # 5 % repeats of the previous syllable (chains are rare; good enough)
    return np.split(ids.astype(np.int64), np.cumsum(L)[:-1])


def planted_logits(batch: ClipBatch, V: int = V_HEAD, ctc: bool = True, device="cuda", seed: int = 114514,
                   scale: float = 2.0, boost: float = 8.0) -> torch.Tensor:
    """[sum T, V] fp32 logits: randn * scale, +boost on the true label column along a random
    monotone segmentation, silence logit raised on blank frames (so alignments are non-trivial)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = int(batch.t_len.sum())
    z = torch.empty((total, V), dtype=torch.float32, device=device)
    step = 1 << 16
    for r in range(0, total, step):                       # chunked: keeps the temporary small
        z[r:r + step].normal_(0.0, scale, generator=g)
    rng = np.random.default_rng(seed + 1)
    rows_l, cols_l, rows_s, val_s = [], [], [], []
    r0 = 0
    for T, lab in zip(batch.t_len, batch.labels):
        T, L = int(T), len(lab)
        ncut = min(2 * L, T - 1)
        cuts = np.sort(rng.choice(np.arange(1, T), size=ncut, replace=False)) if ncut > 0 else np.zeros(0, int)
        seg = np.minimum(np.searchsorted(cuts, np.arange(T), side="right"), 2 * L)
        voiced = (seg % 2 == 1)
        tt = np.nonzero(voiced)[0]
        rows_l.append(r0 + tt)
        cols_l.append(lab[seg[tt] // 2])
        rows_s.append(r0 + np.arange(T))
        val_s.append(np.where(voiced, -3.0, 3.0 if ctc else boost))
        r0 += T
    rows_l = torch.from_numpy(np.concatenate(rows_l)).to(device)
    cols_l = torch.from_numpy(np.concatenate(cols_l)).to(device)
    z[rows_l, cols_l] += boost
    rows_s = torch.from_numpy(np.concatenate(rows_s)).to(device)
    val_s = torch.from_numpy(np.concatenate(val_s).astype(np.float32)).to(device)
    z[rows_s, (V - 1) if ctc else 0] += val_s
    return z


def synthetic_waveforms(batch: ClipBatch, device="cuda", seed: int = 114514):
    """SURVEY.md 8(d) waveforms, concatenated (every clip start padded to a multiple of 4 samples so
    the TMA path applies): 0.1 N(0,1) noise + 10 amplitude-modulated harmonics of 220 Hz, last 10 %
    zeroed. Returns (wave [total] float32 on `device`, offsets int64 [B])."""
    offs = np.zeros(len(batch.n_samples), np.int64)
    pos = 0
    for i, n in enumerate(batch.n_samples):
        offs[i] = pos
        pos += (int(n) + 3) // 4 * 4
    g = torch.Generator(device=device)
    g.manual_seed(seed + 7)
    wave = torch.empty(pos + 8, dtype=torch.float32, device=device)
    wave.normal_(0.0, 0.1, generator=g)
    for i, n in enumerate(batch.n_samples):
        n = int(n)
        t = torch.arange(n, device=device, dtype=torch.float32) / 16000.0
        seg = wave[offs[i]:offs[i] + n]
        env = 0.5 + 0.5 * torch.sin(2 * np.pi * 3.0 * t)
        for h in range(1, 11):
            seg += (0.3 / h) * torch.sin(2 * np.pi * 220.0 * h * t) * env
        seg[int(0.9 * n):] = 0.0
    return wave, offs
