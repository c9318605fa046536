"""N1: the head's ``Linear(768 -> 21129)`` fused with the emission math, so the [T, V] logits are
never materialised.

In the reference the logits are produced by ``self.fc(self.activate(out))`` (module/align_model.py:32-33,38,107),
copied to the host (inference_alignment.py:161: 21-127 MB per clip) and consumed by ``log_softmax`` / ``sigmoid``
(utils/alignment.py:123-134). ``FusedHead`` takes over from the Mish output onwards:

    head = FusedHead(model.align_rnn.fc.weight, model.align_rnn.fc.bias)          # packs the weight once
    hidden = model.align_rnn.activate(model.align_rnn.rnn(embed)[0])              # [B, T, 768], stock PyTorch
    onoff = head.perform_viterbi_ctc(hidden, labels)                             # == perform_viterbi_ctc(fc(hidden), labels)

Hidden states may live on the GPU (zero copies) or in (pinned) host memory: 3 KB per frame cross PCIe instead of
the 84.5 KB of logits. CUDA only, like everything else here.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from . import alignment as A
from ._lib import MODE_CE, MODE_CTC


class FusedHead:
    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device=None):
        A._require_cuda()
        lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self.weight = weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.bias = bias.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.V, self.D = self.weight.shape
        if self.bias.shape != (self.V,):
            raise ValueError("bias must be [V]")
        nbytes = int(lib.la_head_packed_weight_bytes(self.V, self.D))
        if nbytes == 0:
            raise ValueError("hidden width must be a multiple of 32 in [32, 1024]")
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._free_ws: list = []                      # recycled (head workspace, plan workspace) pairs: GBs per batch
        self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.la_head_pack_weights(self.weight.data_ptr(), self.weight.stride(0), self.V, self.D,
                                                self.packed.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                       "la_head_pack_weights")

    # ------------------------------------------------------------------------------------------
    def align_clips_async(self, hidden2d: torch.Tensor, t_len, labels, mode: int = MODE_CTC, timing=None) -> "HeadJob":
        """hidden2d: float32 [sum(t_len), D] (CUDA, or host -- pin it -- which is uploaded first)."""
        if hidden2d.dim() != 2 or hidden2d.shape[1] != self.D or hidden2d.dtype != torch.float32:
            raise ValueError(f"hidden2d must be float32 [frames, {self.D}]")
        lens, flat = A._flatten_labels(labels)
        t_len = np.ascontiguousarray(t_len, dtype=np.int32)
        if len(lens) != len(t_len) or int(t_len.sum()) != hidden2d.shape[0]:
            raise ValueError("t_len / labels do not match the hidden states")
        l_len, cols = A._resolve_columns((lens, flat), self.V - 2 if mode == MODE_CTC else self.V - 1)
        with torch.cuda.device(self.device):
            if hidden2d.is_cuda:
                x = hidden2d
            else:
                # host hidden states (pin them): the upload runs on a copy stream, so with two jobs in flight the
                # copy of batch i+1 overlaps the GEMM of batch i; 3 KB per frame instead of 84.5 KB of logits
                cur = torch.cuda.current_stream(self.device)
                with torch.cuda.stream(self._copy_stream):
                    x = hidden2d.to(self.device, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(self._copy_stream)
                cur.wait_event(done)
                x.record_stream(cur)
            if x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16:
                x = x.contiguous().clone()
            plan = A.AlignPlan(mode, self.V, t_len, l_len, cols, self.device.index)
            try:
                return HeadJob(self, plan, x, timing)
            except Exception:
                plan.close()
                raise

    def align_clips(self, hidden2d, t_len, labels, mode: int = MODE_CTC) -> A.AlignResult:
        return self.align_clips_async(hidden2d, t_len, labels, mode).result()

    def align(self, hidden: torch.Tensor, labels, mode: int = MODE_CTC) -> A.AlignResult:
        """hidden: [B, T, D] padded batch, like the ``prediction`` of perform_viterbi*."""
        if hidden.dim() != 3:
            raise ValueError("hidden must be [batch, frames, width]")
        B, T, D = hidden.shape
        return self.align_clips(hidden.detach().float().contiguous().view(B * T, D), np.full(B, T, np.int32),
                                self._rows(labels, B), mode)

    @staticmethod
    def _rows(labels, B):
        lens, flat = A._flatten_labels(labels)
        if len(lens) < B:
            raise IndexError("fewer label rows than batch items")
        return np.split(flat[:int(lens[:B].sum())], np.cumsum(lens[:B])[:-1]) if B else []

    def perform_viterbi_ctc(self, hidden, labels, hop_size_second: float = 0.02):
        """== alignment.perform_viterbi_ctc(fc(hidden), labels) without the logits (utils/alignment.py:121-188)."""
        return A.onoff_seconds(self.align(hidden, labels, MODE_CTC), hop_size_second)

    def perform_viterbi(self, hidden, labels, hop_size_second: float = 0.02):
        """== alignment.perform_viterbi(fc(hidden), labels) (utils/alignment.py:13-71)."""
        return A.onoff_seconds(self.align(hidden, labels, MODE_CE), hop_size_second)


class HeadJob:
    """la_head_align enqueued on the current stream; ``result()`` waits for this job only (see AlignJob)."""

    def __init__(self, head: FusedHead, plan: A.AlignPlan, x: torch.Tensor, timing=None):
        lib = _lib.load()
        dev = x.device
        self.plan, self._x, self._head = plan, x, head
        need_ws, need_hws = plan.workspace_bytes, int(lib.la_head_workspace_bytes(plan.handle, head.D))
        self.ws = self.head_ws = None
        for i, (hws, ws) in enumerate(head._free_ws):               # stream-ordered reuse: same stream, so no hazard
            if hws.numel() >= need_hws and ws.numel() >= need_ws:
                self.head_ws, self.ws = head._free_ws.pop(i)
                break
        if self.ws is None:
            self.ws = torch.empty(need_ws, dtype=torch.uint8, device=dev)
            self.head_ws = torch.empty(need_hws, dtype=torch.uint8, device=dev)
        B, Ltot = max(plan.n_utt, 1), max(plan.total_labels, 1)
        self._B, self._Ltot = B, Ltot
        nbytes = 8 * B + 4 * (2 * Ltot + B)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        score = packed[:8 * B].view(torch.float64)
        ints = packed[8 * B:].view(torch.int32)
        first, last, status = ints[:Ltot], ints[Ltot:2 * Ltot], ints[2 * Ltot:2 * Ltot + B]
        stream = torch.cuda.current_stream(dev)
        if timing is not None:
            timing[0].record(stream)
        _lib.check(lib.la_head_emit(plan.handle, x.data_ptr(), x.stride(0), head.D, head.weight.data_ptr(),
                                    head.weight.stride(0), head.bias.data_ptr(), head.packed.data_ptr(),
                                    self.head_ws.data_ptr(), self.ws.data_ptr(), stream.cuda_stream), "la_head_emit")
        if timing is not None:
            timing[1].record(stream)
        _lib.check(lib.la_viterbi(plan.handle, self.ws.data_ptr(), first.data_ptr(), last.data_ptr(),
                                  score.data_ptr(), status.data_ptr(), stream.cuda_stream), "la_viterbi")
        self._pinned = A._pinned_take(nbytes)
        self._pinned[:nbytes].copy_(packed, non_blocking=True)
        self._packed, self._nbytes = packed, nbytes
        self._event = torch.cuda.Event()
        self._event.record(stream)
        self._res = None

    def result(self) -> A.AlignResult:
        if self._res is None:
            self._event.synchronize()
            plan, B, Ltot = self.plan, self._B, self._Ltot
            host = self._pinned[:self._nbytes].numpy().copy()
            A._pinned_give(self._pinned)
            h_score = host[:8 * B].view(np.float64)
            h_int = host[8 * B:].view(np.int32)
            self._res = A.AlignResult(h_int[:plan.total_labels], h_int[Ltot:Ltot + plan.total_labels],
                                      h_score[:plan.n_utt], h_int[2 * Ltot:2 * Ltot + plan.n_utt], plan.l_len)
            self._pinned = self._packed = self._x = None
        return self._res

    def close(self):
        if self.ws is not None and len(self._head._free_ws) < 4:
            self._head._free_ws.append((self.head_ws, self.ws))
        self.ws = self.head_ws = None
        self.plan.close()
