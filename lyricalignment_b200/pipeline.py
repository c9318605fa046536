"""End-to-end stand-in for ``AlignModel.frame_manual_forward`` (module/align_model.py:72-123) with
the product's kernels either side of the STOCK encoder + head.

The Whisper encoder and the GRU / Mish / Linear head are outside the product (north_star): they stay
plain PyTorch modules here -- ``transformers``' WhisperEncoder with random weights stands in for
``whisper_model.embed_audio`` (no checkpoints offline) and ``AlignHead`` restates the reference's
``RNN`` (module/align_model.py:11-40). What this module adds is the framing arithmetic of the
reference around them, on the device:

  * audios zero-padded to the batch maximum ON THE DEVICE (align_model.py:78-82), ONE log-mel call with the
    global max (K1, :84) written directly into the zero-padded encoder window(s),
  * <= 3000 mel frames: T = int(round(F / 2.0)) (half-to-even), encoder on the full 30 s window, output
    sliced to T (:87-92),
  * > 3000 frames: independent 3000-frame chunks, each sliced to round(len / 2), concatenated (:93-104),
  * head over the concatenated embedding -> logits [B, T, V] on the device, ready for
    ``perform_viterbi_ctc`` without the ``.cpu()`` round trip.

Used by scripts/bench_end_to_end.py to show the decode path's share of wall time and
by tests/test_gpu_pipeline.py.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
from torch import nn

from . import _lib
from . import audio as LA

WHISPER_DIMS = {"tiny": (384, 4, 6), "base": (512, 6, 8), "small": (768, 12, 12),
                "medium": (1024, 24, 16), "large": (1280, 32, 20)}     # d_model, layers, heads


class AlignHead(nn.Module):
    """The reference's RNN head: GRU(2 layers, bidirectional, hidden 384) -> Mish -> Linear(768 -> V)."""

    def __init__(self, input_size: int, hidden_size: int = 384, output_size: int = 21129,
                 num_layers: int = 2, dropout: float = 0.1, bidirectional: bool = True):
        super().__init__()
        self.rnn = nn.GRU(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers,
                          dropout=dropout, batch_first=True, bidirectional=bidirectional)
        self.activate = nn.Mish()
        self.fc = nn.Linear(hidden_size * (2 if bidirectional else 1), output_size)

    def forward(self, x):
        out, _ = self.rnn(x)
        return self.fc(self.activate(out))


def make_encoder(size: str = "tiny") -> nn.Module:
    """Random-init stock encoder with Whisper's architecture (2 convs + n pre-LN blocks, 1500 positions)."""
    from transformers import WhisperConfig
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    d, layers, heads = WHISPER_DIMS[size]
    cfg = WhisperConfig(d_model=d, encoder_layers=layers, encoder_attention_heads=heads, encoder_ffn_dim=4 * d,
                        num_mel_bins=80, max_source_positions=1500)
    return WhisperEncoder(cfg)


class AlignPipeline(nn.Module):
    def __init__(self, size: str = "tiny", vocab: int = 21129, device="cuda", seed: int = 114514):
        super().__init__()
        torch.manual_seed(seed)
        self.encoder = make_encoder(size)
        self.head = AlignHead(WHISPER_DIMS[size][0], 384, vocab)
        self.to(device).eval()
        self.device = torch.device(device)

    def embed_audio(self, mel: torch.Tensor) -> torch.Tensor:
        return self.encoder(mel).last_hidden_state

    @torch.no_grad()
    def frame_manual_forward(self, audios: Sequence[np.ndarray]) -> torch.Tensor:
        """module/align_model.py:72-123 with every byte of the framing on the device:
        one pinned concat + ONE H2D of the ragged clips, the zero-padded [B, n_max] batch built by a single
        masked scatter (align_model.py:78-82 does it with np.append per item on the host), and K1 writing the
        log-mel straight into a pre-zeroed [B, 80, 3000 * n_chunks] tensor -- which IS the pad_or_trim'ed
        encoder window of every chunk (align_model.py:89,100; 0.0 is the pad value in normalised log-mel
        units), so no pad_or_trim copy is made."""
        lib = _lib.load()
        dev = self.device
        lens = np.fromiter((len(a) for a in audios), dtype=np.int64, count=len(audios))
        B, n = len(audios), int(lens.max())
        flat = torch.from_numpy(np.concatenate([np.asarray(a, dtype=np.float32).reshape(-1) for a in audios]))
        flat = flat.pin_memory().to(dev, non_blocking=True)
        lens_d = torch.from_numpy(lens).to(dev, non_blocking=True)
        n_pad = (n + 3) // 4 * 4                                        # rows 16-byte aligned -> K1's TMA path
        batch = torch.zeros((B, n_pad), dtype=torch.float32, device=dev)
        batch[torch.arange(n_pad, device=dev)[None, :] < lens_d[:, None]] = flat      # align_model.py:78-82
        F_ = n // LA.HOP_LENGTH
        n_chunks = max(1, -(-F_ // LA.N_FRAMES))
        mel = torch.zeros((B, LA.N_MELS, LA.N_FRAMES * n_chunks), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = torch.empty(int(lib.la_logmel_workspace_bytes(B, B * n)), dtype=torch.uint8, device=dev)
            _lib.check(lib.la_logmel(batch.data_ptr(), B, n, n_pad, mel.data_ptr(), mel.shape[-1], ws.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), "la_logmel")   # K1, global max (:84)
        if F_ <= LA.N_FRAMES:
            T = LA.decode_frames(F_)                                   # :88
            embed = self.embed_audio(mel)[:, :T, :]
        else:
            parts = []
            for s in range(0, F_, LA.N_FRAMES):                        # :95-104
                e = min(s + LA.N_FRAMES, F_)
                parts.append(self.embed_audio(mel[:, :, s:s + LA.N_FRAMES])[:, :LA.decode_frames(e - s), :])
            embed = torch.cat(parts, dim=1)
        return self.head(embed)                                        # [B, T, V] on the device
