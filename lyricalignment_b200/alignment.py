"""Drop-in for the reference's ``utils/alignment.py`` (navi0105/LyricAlignment), running on
hand-written sm_100a kernels through the C ABI in include/lyricalign.h.

Same names, argument meaning, return types and exceptions as the reference:

  perform_viterbi_ctc(prediction, labels, hop_size_second=0.02)   utils/alignment.py:121-188
  perform_viterbi(prediction, labels, hop_size_second=0.02)       utils/alignment.py:13-71
  run_viterbi_core(dp, bt, logp, sil, label)                      utils/alignment.py:73-119
  get_mae(gt, predict)                                            utils/alignment.py:190-199

``prediction`` may be a CUDA tensor (fast path: the entry scripts simply drop their ``.cpu()``,
inference_alignment.py:161) or a CPU tensor / ndarray (the reference's actual call; it is streamed
to the GPU through the library's double-buffered host path). There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MODE_CE, MODE_CTC, MODE_LOGP

__all__ = ["perform_viterbi_ctc", "perform_viterbi", "run_viterbi_core", "get_mae", "align",
           "align_clips", "align_clips_async", "AlignJob", "AlignResult", "AlignPlan"]


# ----------------------------------------------------------------------------------------------
# labels
# ----------------------------------------------------------------------------------------------
def _label_rows(labels) -> List[np.ndarray]:
    """utils/alignment.py:141: keep everything that is not the -100 padding."""
    lens, flat = _flatten_labels(labels)
    out, p = [], 0
    for n in lens.tolist():
        out.append(flat[p:p + n])
        p += n
    return out


def _flatten_labels(labels):
    """-> (lens int32 [B], flat int64 [sum lens]) with the -100 padding stripped (utils/alignment.py:141),
    in one vectorised pass: the reference's per-element tensor loop costs ~1 ms per clip; a per-row numpy
    loop would still cost milliseconds per 2 000-clip batch once the kernels take 16 ms."""
    if torch.is_tensor(labels):
        labels = labels.detach().cpu().numpy()
    if isinstance(labels, np.ndarray) and labels.ndim == 2:
        lab = labels.astype(np.int64, copy=False)
        keep = lab != -100
        return keep.sum(axis=1).astype(np.int32), lab[keep]
    rows = [r if (isinstance(r, np.ndarray) and r.dtype == np.int64 and r.ndim == 1)
            else np.asarray(r, dtype=np.int64).reshape(-1) for r in labels]
    if not rows:
        return np.zeros(0, np.int32), np.zeros(0, np.int64)
    lens = np.fromiter((r.size for r in rows), dtype=np.int64, count=len(rows))
    flat = np.concatenate(rows) if int(lens.sum()) else np.zeros(0, np.int64)
    keep = flat != -100
    if not keep.all():
        row_of = np.repeat(np.arange(len(rows)), lens)
        lens = np.bincount(row_of[keep], minlength=len(rows))
        flat = flat[keep]
    return lens.astype(np.int32), flat


def _resolve_columns(rows, ncols: int):
    """Label c indexes column c-1 of the reference's sliced emission matrix (ncols wide), i.e.
    ORIGINAL logit column (c-1)+1. numpy/numba wrap a negative index once (c <= 0); anything else
    out of range would be an out-of-bounds read in the reference's nopython kernel -- refused.
    `rows`: a list of label rows, or the (lens, flat) pair of _flatten_labels."""
    if isinstance(rows, tuple):
        lens, flat = rows
    else:
        lens = np.array([len(r) for r in rows], dtype=np.int32)
        flat = np.concatenate(rows) if len(rows) and lens.sum() else np.zeros(0, np.int64)
    col = flat - 1
    col = np.where(col < 0, col + ncols, col)
    if col.size and (col.min() < 0 or col.max() >= ncols):
        raise IndexError(f"label id outside the emission matrix (valid ids: 1..{ncols})")
    return lens, (col + 1).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# plan + raw alignment
# ----------------------------------------------------------------------------------------------
class AlignPlan:
    """Owns a la_plan: the ragged shapes and resolved label columns of one batch."""

    def __init__(self, mode: int, V: int, t_len, l_len, columns, device: int):
        lib = _lib.load()
        self.t_len = np.ascontiguousarray(t_len, dtype=np.int32)
        self.l_len = np.ascontiguousarray(l_len, dtype=np.int32)
        self.columns = np.ascontiguousarray(columns, dtype=np.int32)
        assert len(self.t_len) == len(self.l_len) and self.columns.size == int(self.l_len.sum())
        self.mode, self.V, self.device = mode, V, device
        self._h = ctypes.c_void_p()
        _lib.check(lib.la_plan_create(ctypes.byref(self._h), mode, len(self.t_len), V,
                                      self.t_len.ctypes.data, self.l_len.ctypes.data,
                                      self.columns.ctypes.data if self.columns.size else None, device),
                   "la_plan_create")
        self.n_utt = len(self.t_len)
        self.total_frames = int(lib.la_plan_total_frames(self._h))
        self.total_labels = int(lib.la_plan_total_labels(self._h))
        self.workspace_bytes = int(lib.la_plan_workspace_bytes(self._h))
        self.num_launches = int(lib.la_plan_num_launches(self._h))

    @property
    def handle(self):
        return self._h

    def utt_layout(self, u: int):
        eo, rf, bo, pp = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int32()
        _lib.check(_lib.load().la_plan_utt_layout(self._h, u, ctypes.byref(eo), ctypes.byref(rf),
                                                  ctypes.byref(bo), ctypes.byref(pp)), "la_plan_utt_layout")
        return eo.value, rf.value, bo.value, pp.value

    def close(self):
        if self._h:
            _lib.load().la_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class AlignResult:
    first: np.ndarray        # int32 [sum L]   onset frame of every label
    last_plus1: np.ndarray   # int32 [sum L]   offset frame (exclusive)
    score: np.ndarray        # float64 [B]     dp[T-1][end state]
    status: np.ndarray       # int32 [B]       0 ok / 1 empty labels / 2 infeasible
    l_len: np.ndarray        # int32 [B]


_cuda_ok = None


def _require_cuda():
    global _cuda_ok
    if _cuda_ok is None:
        _cuda_ok = bool(torch.cuda.is_available())
    if not _cuda_ok:
        raise _lib.LyricAlignError("lyricalignment_b200 needs a CUDA device (no CPU fallback)")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_pinned_pool: dict = {}          # size class -> [pinned uint8 tensors]; cudaHostAlloc costs ~100 us, a 2 000-clip batch 16 ms


def _pinned_take(nbytes: int) -> torch.Tensor:
    cls = 1 << max(12, (nbytes - 1).bit_length())
    free = _pinned_pool.setdefault(cls, [])
    return free.pop() if free else torch.empty(cls, dtype=torch.uint8, pin_memory=True)


def _pinned_give(buf: torch.Tensor) -> None:
    free = _pinned_pool.setdefault(buf.numel(), [])
    if len(free) < 8:
        free.append(buf)


class AlignJob:
    """K2 + K3 enqueued on the current stream, results on their way to a pinned host buffer.
    ``result()`` waits for THIS job only (an event), so the host can prepare the next batch -- label
    flattening, plan creation -- while the kernels of this one run. Owns the plan until then."""

    def __init__(self, plan: AlignPlan, logits2d: torch.Tensor, sil: torch.Tensor | None = None,
                 keep_workspace: bool = False, timing=None):
        """timing: optional (start, end) torch.cuda.Event pair recorded around K2 (the benchmark's roofline leg);
        K2 and K3 are then enqueued as la_emit + la_viterbi instead of la_align -- the same two kernels."""
        lib = _lib.load()
        dev = logits2d.device
        assert logits2d.dtype == torch.float32 and logits2d.stride(1) == 1
        if logits2d.data_ptr() % 16:
            logits2d = logits2d.clone()
        ld = logits2d.stride(0) if logits2d.shape[0] > 1 else max(logits2d.stride(0), plan.V)
        self.plan = plan
        self._logits = logits2d                       # keep the inputs alive while the kernels run
        self.ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
        # one packed result buffer -> one D2H: [score f64 x B | first i32 x L | last i32 x L | status i32 x B]
        B, Ltot = max(plan.n_utt, 1), max(plan.total_labels, 1)
        self._B, self._Ltot = B, Ltot
        nbytes = 8 * B + 4 * (2 * Ltot + B)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        score = packed[:8 * B].view(torch.float64)
        ints = packed[8 * B:].view(torch.int32)
        first, last, status = ints[:Ltot], ints[Ltot:2 * Ltot], ints[2 * Ltot:2 * Ltot + B]
        stream = _stream_ptr(dev)
        if plan.mode == MODE_LOGP or timing is not None:
            if timing is not None:
                timing[0].record(torch.cuda.current_stream(dev))
            _lib.check(lib.la_emit(plan.handle, logits2d.data_ptr(), ld, sil.data_ptr() if sil is not None else None,
                                   sil.stride(0) if sil is not None else 0, self.ws.data_ptr(), stream), "la_emit")
            if timing is not None:
                timing[1].record(torch.cuda.current_stream(dev))
            _lib.check(lib.la_viterbi(plan.handle, self.ws.data_ptr(), first.data_ptr(), last.data_ptr(),
                                      score.data_ptr(), status.data_ptr(), stream), "la_viterbi")
        else:
            _lib.check(lib.la_align(plan.handle, logits2d.data_ptr(), ld, self.ws.data_ptr(), first.data_ptr(),
                                    last.data_ptr(), score.data_ptr(), status.data_ptr(), stream), "la_align")
        self._pinned = _pinned_take(nbytes)
        self._pinned[:nbytes].copy_(packed, non_blocking=True)
        self._packed = packed
        self._nbytes = nbytes
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(dev))
        self._keep_ws = keep_workspace
        self._res = None

    def done(self) -> bool:
        return self._res is not None or self._event.query()

    def result(self) -> AlignResult:
        if self._res is None:
            self._event.synchronize()
            plan, B, Ltot = self.plan, self._B, self._Ltot
            host = self._pinned[:self._nbytes].numpy().copy()      # the pinned buffer goes back to the pool
            _pinned_give(self._pinned)
            h_score = host[:8 * B].view(np.float64)
            h_int = host[8 * B:].view(np.int32)
            self._res = AlignResult(h_int[:plan.total_labels], h_int[Ltot:Ltot + plan.total_labels],
                                    h_score[:plan.n_utt], h_int[2 * Ltot:2 * Ltot + plan.n_utt], plan.l_len)
            self._pinned = self._packed = self._logits = None
            if not self._keep_ws:
                self.ws = None
                plan.close()
        return self._res


def _run_device(plan: AlignPlan, logits2d: torch.Tensor, sil: torch.Tensor | None = None,
                keep_workspace: bool = False):
    """logits2d: CUDA float32 [sum T, V] with unit column stride."""
    job = AlignJob(plan, logits2d, sil, keep_workspace=True)
    res = job.result()
    return (res, job.ws) if keep_workspace else res


def _run_host(plan: AlignPlan, logits2d: torch.Tensor, staging_bytes: int = 0) -> AlignResult:
    """logits2d: CPU float32 [sum T, V] contiguous (pinned for full PCIe rate)."""
    lib = _lib.load()
    assert logits2d.dtype == torch.float32 and logits2d.is_contiguous()
    first = np.empty(max(plan.total_labels, 1), np.int32)
    last = np.empty(max(plan.total_labels, 1), np.int32)
    score = np.empty(max(plan.n_utt, 1), np.float64)
    status = np.empty(max(plan.n_utt, 1), np.int32)
    _lib.check(lib.la_align_host(plan.handle, logits2d.data_ptr(), logits2d.shape[1], first.ctypes.data,
                                 last.ctypes.data, score.ctypes.data, status.ctypes.data, staging_bytes),
               "la_align_host")
    return AlignResult(first[:plan.total_labels], last[:plan.total_labels], score[:plan.n_utt],
                       status[:plan.n_utt], plan.l_len)


def align(prediction, labels, mode: int = MODE_CTC, device: int | None = None) -> AlignResult:
    """Raw frame-index alignment of a padded batch ``prediction`` [B, T, V] against ``labels``."""
    _require_cuda()
    if not torch.is_tensor(prediction):
        prediction = torch.from_numpy(np.ascontiguousarray(prediction, dtype=np.float32))
    if prediction.dim() != 3:
        raise ValueError("prediction must be [batch, frames, vocab]")
    prediction = prediction.detach()
    if prediction.dtype != torch.float32:
        prediction = prediction.float()
    B, T, V = prediction.shape
    lens, flat = _flatten_labels(labels)
    if len(lens) < B:
        raise IndexError("fewer label rows than batch items")
    if len(lens) > B:
        flat = flat[:int(lens[:B].sum())]
        lens = lens[:B]
    l_len, cols = _resolve_columns((lens, flat), V - 2 if mode == MODE_CTC else V - 1)
    if prediction.is_cuda:
        dev = prediction.device.index
        with torch.cuda.device(dev):
            plan = AlignPlan(mode, V, np.full(B, T, np.int32), l_len, cols, dev)
            try:
                return _run_device(plan, prediction.contiguous().view(B * T, V))
            finally:
                plan.close()
    dev = torch.cuda.current_device() if device is None else device
    plan = AlignPlan(mode, V, np.full(B, T, np.int32), l_len, cols, dev)
    try:
        return _run_host(plan, prediction.contiguous().view(B * T, V))
    finally:
        plan.close()


def align_clips(logits2d: torch.Tensor, t_len, labels, mode: int = MODE_CTC, device: int | None = None,
                staging_bytes: int = 0) -> AlignResult:
    """Ragged batch in one call: ``logits2d`` is float32 [sum(t_len), V] with clip u owning rows
    [sum(t_len[:u]), +t_len[u]); ``labels`` one row of class ids per clip. CUDA tensor -> K2 + K3 in
    place; CPU tensor (pin it) -> streamed through the double-buffered host path, so the H2D copy
    of clip u+1 overlaps the kernels of clip u. This is the call to use for a whole dataset; the
    reference-shaped ``perform_viterbi*`` wrappers go through the same code with B padded clips."""
    _require_cuda()
    if logits2d.dim() != 2 or logits2d.dtype != torch.float32:
        raise ValueError("logits2d must be float32 [frames, vocab]")
    V = logits2d.shape[1]
    lens, flat = _flatten_labels(labels)
    t_len = np.ascontiguousarray(t_len, dtype=np.int32)
    if len(lens) != len(t_len) or int(t_len.sum()) != logits2d.shape[0]:
        raise ValueError("t_len / labels do not match the logits")
    l_len, cols = _resolve_columns((lens, flat), V - 2 if mode == MODE_CTC else V - 1)
    if logits2d.is_cuda:
        dev = logits2d.device.index
        with torch.cuda.device(dev):
            plan = AlignPlan(mode, V, t_len, l_len, cols, dev)
            try:
                return _run_device(plan, logits2d if logits2d.stride(1) == 1 else logits2d.contiguous())
            finally:
                plan.close()
    dev = torch.cuda.current_device() if device is None else device
    plan = AlignPlan(mode, V, t_len, l_len, cols, dev)
    try:
        return _run_host(plan, logits2d.contiguous(), staging_bytes)
    finally:
        plan.close()


def align_clips_async(logits2d: torch.Tensor, t_len, labels, mode: int = MODE_CTC, timing=None) -> AlignJob:
    """``align_clips`` for CUDA logits without the wait: enqueues K2 + K3 + the result copy on the current
    stream and returns at once. Call ``.result()`` when the alignment is needed; keep one or two jobs in
    flight and the host-side work of the next batch (labels, plan) hides behind the kernels of this one."""
    _require_cuda()
    if logits2d.dim() != 2 or logits2d.dtype != torch.float32 or not logits2d.is_cuda:
        raise ValueError("logits2d must be CUDA float32 [frames, vocab]")
    V = logits2d.shape[1]
    lens, flat = _flatten_labels(labels)
    t_len = np.ascontiguousarray(t_len, dtype=np.int32)
    if len(lens) != len(t_len) or int(t_len.sum()) != logits2d.shape[0]:
        raise ValueError("t_len / labels do not match the logits")
    l_len, cols = _resolve_columns((lens, flat), V - 2 if mode == MODE_CTC else V - 1)
    dev = logits2d.device.index
    with torch.cuda.device(dev):
        plan = AlignPlan(mode, V, t_len, l_len, cols, dev)
        try:
            return AlignJob(plan, logits2d if logits2d.stride(1) == 1 else logits2d.contiguous(), timing=timing)
        except Exception:
            plan.close()
            raise


def onoff_seconds(res: AlignResult, hop_size_second: float = 0.02):
    """AlignResult -> the reference's nested [[onset, offset], ...] lists (raises like it does)."""
    return _to_onoff(res, hop_size_second)


def _to_onoff(res: AlignResult, hop_size_second: float) -> List[List[List[float]]]:
    status = res.status
    if status.size and status.max() != 0:
        u = int(np.nonzero(status)[0][0])          # first failing utterance in batch order decides
        if int(status[u]) == _lib.UTT_EMPTY:        # reference: cur_label[0] on an empty array (:152)
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")
        raise ValueError("label state is not in list")   # reference: correct_path.index(k * 2 + 1) (:183)
    # utils/alignment.py:185: float(index) * hop in Python fp64 == IEEE fp64 multiply, done in one numpy pass
    hop = float(hop_size_second)
    pairs = np.stack([res.first.astype(np.float64) * hop, res.last_plus1.astype(np.float64) * hop], axis=1).tolist()
    out, p = [], 0
    for n in res.l_len.tolist():
        out.append(pairs[p:p + n])
        p += n
    return out


def perform_viterbi_ctc(prediction, labels, hop_size_second: float = 0.02):
    """CTC-trained models: softmax over columns 1..V-2, sigmoid silence in column V-1."""
    return _to_onoff(align(prediction, labels, MODE_CTC), hop_size_second)


def perform_viterbi(prediction, labels, hop_size_second: float = 0.02):
    """CE-trained models (BASELINE.json's "DTW" config): softmax over all V, silence = column 0."""
    return _to_onoff(align(prediction, labels, MODE_CE), hop_size_second)


def run_viterbi_core(dp_matrix, backtrace_dp_matrix, cur_log_prediction, cur_log_silence_prediction,
                     cur_label):
    """The reference's inner boundary (utils/alignment.py:73-119), kept for parity work: same
    arguments, fills ``dp_matrix`` (fp64) and ``backtrace_dp_matrix`` (int64) in place for rows
    1..T-1 and returns them. The DP runs on the GPU from the caller's log-prob matrices; the
    production kernel never materialises these tables (2-bit step codes only), so this goes
    through the library's parity instrumentation (la_viterbi_debug). Row 0 is the caller's preset,
    as in the reference -- it must be the standard one (:144-152)."""
    _require_cuda()
    lib = _lib.load()
    logp = torch.as_tensor(np.ascontiguousarray(cur_log_prediction, dtype=np.float32)).cuda()
    sil = torch.as_tensor(np.ascontiguousarray(cur_log_silence_prediction, dtype=np.float32).reshape(-1)).cuda()
    lab = np.asarray(cur_label, dtype=np.int64)
    T, ncols = logp.shape
    S = 2 * len(lab) + 1
    l_len, cols = _resolve_columns([lab], ncols)
    dev = logp.device
    plan = AlignPlan(MODE_LOGP, ncols, np.array([T], np.int32), l_len, cols - 1, dev.index)
    try:
        ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
        first = torch.empty(len(lab), dtype=torch.int32, device=dev)
        last = torch.empty(len(lab), dtype=torch.int32, device=dev)
        score = torch.empty(1, dtype=torch.float64, device=dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        dp = torch.empty(T * S, dtype=torch.float64, device=dev)
        stream = _stream_ptr(dev)
        _lib.check(lib.la_emit(plan.handle, logp.data_ptr(), logp.stride(0), sil.data_ptr(), 1,
                               ws.data_ptr(), stream), "la_emit")
        _lib.check(lib.la_viterbi_debug(plan.handle, ws.data_ptr(), first.data_ptr(), last.data_ptr(),
                                        score.data_ptr(), status.data_ptr(), dp.data_ptr(), stream),
                   "la_viterbi_debug")
        codes = unpack_step_codes(plan, ws, 0)
        dp_h = dp.cpu().numpy().reshape(T, S)
    finally:
        plan.close()
    dp_matrix[1:, :] = dp_h[1:]
    backtrace_dp_matrix[1:, :] = np.arange(S)[None, :] - codes[1:]
    return dp_matrix, backtrace_dp_matrix


def unpack_step_codes(plan: AlignPlan, ws: torch.Tensor, u: int) -> np.ndarray:
    """Expands utterance u's packed backpointers into codes[T][2L+1] (k - bt[t][k]; row 0 unused)."""
    eo, rf, bo, pp = plan.utt_layout(u)
    nrow, shift = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.load().la_plan_utt_bp_layout(plan.handle, u, ctypes.byref(nrow), ctypes.byref(shift)),
               "la_plan_utt_bp_layout")
    T, L = int(plan.t_len[u]), int(plan.l_len[u])
    nblk, sh = nrow.value, shift.value
    words = ws[bo:bo + nblk * pp * 4].cpu().numpy().view(np.uint32).reshape(nblk, pp)
    t = np.arange(T)
    nib = (words[t // 8][:, sh:sh + L + 1] >> ((t % 8) * 4)[:, None].astype(np.uint32)) & 0xF    # [T][L+1]
    codes = np.zeros((T, 2 * L + 1), np.int64)
    codes[:, 0::2] = nib & 1
    codes[:, 1::2] = (nib >> 1)[:, :L]
    return codes


def unpack_emissions(plan: AlignPlan, ws: torch.Tensor, u: int) -> np.ndarray:
    """Utterance u's compact emission rows [T][1+L] (column 0 = blank)."""
    eo, rf, bo, pp = plan.utt_layout(u)
    T, L = int(plan.t_len[u]), int(plan.l_len[u])
    e = ws[eo:eo + T * rf * 4].cpu().numpy().view(np.float32).reshape(T, rf)
    return e[:, :L + 1].copy()


def get_mae(gt, predict) -> float:
    """Mean absolute on/offset error over a batch, accumulated in Python fp64 in the
    reference's order (utils/alignment.py:190-199)."""
    total, n = 0.0, 0
    for i in range(len(gt)):
        for j in range(len(gt[i])):
            total = total + abs(gt[i][j][0] - predict[i][j][0]) + abs(gt[i][j][1] - predict[i][j][1])
            n += 2
    return total / float(n)
