"""oracle/viterbi.py -- TEST INFRASTRUCTURE ONLY (CPU checker, never the product path).

ctypes front for oracle/viterbi_oracle.c plus a restatement of the reference's two decoder
entry points on top of it (utils/alignment.py:13-71 ``perform_viterbi``, :121-188
``perform_viterbi_ctc``, :190-199 ``get_mae``), with the reference's return types and
exception behaviour (IndexError on an empty label row, ValueError when a label state is not
on the path).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

from .emission import emission_ce, emission_ctc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "viterbi_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (seconds). Called by __graft_entry__.build() and lazily."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(_OUT_DIR, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", _SO, _SRC])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        lib.la_oracle_align.restype = ctypes.c_int
        lib.la_oracle_align.argtypes = [vp, i64, i64, vp, i64, vp, i64, i64, vp, vp, vp, vp, vp, vp]
        lib.la_oracle_viterbi_core.restype = ctypes.c_int
        lib.la_oracle_viterbi_core.argtypes = [vp, vp, vp, i64, i64, vp, i64, vp, i64, i64]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def viterbi_core(dp, bt, logp, sil, label):
    """Same contract as the reference's numba ``run_viterbi_core`` (utils/alignment.py:73-119):
    in place on dp f64[T,S] / bt i64[T,S]; logp f32[T,V'] (row-strided ok), sil f32[T,1]."""
    assert dp.dtype == np.float64 and bt.dtype == np.int64 and dp.flags.c_contiguous and bt.flags.c_contiguous
    assert logp.dtype == np.float32 and sil.dtype == np.float32
    assert logp.strides[1] == 4 and logp.strides[0] % 4 == 0 and sil.strides[0] % 4 == 0
    label = np.ascontiguousarray(label, dtype=np.int64)
    rc = _load().la_oracle_viterbi_core(_ptr(dp), _ptr(bt), _ptr(logp), logp.strides[0] // 4,
                                        logp.shape[1], _ptr(sil), sil.strides[0] // 4,
                                        _ptr(label), len(label), logp.shape[0])
    if rc:
        raise IndexError(f"label column out of range (status {rc})")
    return dp, bt


def align_one(logp, sil, label, want_tables: bool = False):
    """One utterance: emissions -> dict(status, path, first, last_plus1, score[, dp, bt])."""
    logp = np.asarray(logp)
    sil = np.asarray(sil)
    assert logp.dtype == np.float32 and sil.dtype == np.float32
    assert logp.strides[1] == 4 and logp.strides[0] % 4 == 0
    if sil.ndim == 1:
        sil = sil[:, None]
    label = np.ascontiguousarray(label, dtype=np.int64)
    T, L = logp.shape[0], len(label)
    S = 2 * L + 1
    path = np.zeros(T, np.int32)
    first = np.zeros(L, np.int32)
    last = np.zeros(L, np.int32)
    score = np.zeros(1, np.float64)
    dp = np.empty((T, S), np.float64) if want_tables else None
    bt = np.empty((T, S), np.int64) if want_tables else None
    rc = _load().la_oracle_align(_ptr(logp), logp.strides[0] // 4, logp.shape[1], _ptr(sil),
                                 sil.strides[0] // 4, _ptr(label), L, T, _ptr(path), _ptr(first),
                                 _ptr(last), _ptr(score), _ptr(dp), _ptr(bt))
    out = dict(status=rc, path=path, first=first, last_plus1=last, score=float(score[0]))
    if want_tables:
        out["dp"], out["bt"] = dp, bt
    return out


def _strip(labels_row) -> np.ndarray:
    return np.array([int(x) for x in labels_row if int(x) != -100], dtype=np.int64)   # :141


def _decode(emit, blank, labels, hop_size_second) -> List[List[List[float]]]:
    out = []
    for i in range(emit.shape[0]):                                                     # :140
        lab = _strip(labels[i])
        r = align_one(emit[i], blank[i], lab)
        if r["status"] == 1:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")        # :152
        if r["status"] == 2:
            raise ValueError("label state is not in list")                             # :183
        if r["status"]:
            raise IndexError(f"label column out of range (status {r['status']})")
        out.append([[float(int(f)) * hop_size_second, float(int(l)) * hop_size_second]  # :185
                    for f, l in zip(r["first"], r["last_plus1"])])
    return out


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def perform_viterbi_ctc(prediction, labels: Sequence, hop_size_second: float = 0.02):
    """utils/alignment.py:121-188."""
    emit, blank = emission_ctc(_np(prediction).astype(np.float32, copy=False))
    return _decode(emit, blank, _np(labels) if hasattr(labels, "detach") else labels, hop_size_second)


def perform_viterbi(prediction, labels: Sequence, hop_size_second: float = 0.02):
    """utils/alignment.py:13-71 (the CE-trained decoder; BASELINE.json's "DTW" config)."""
    emit, blank = emission_ce(_np(prediction).astype(np.float32, copy=False))
    return _decode(emit, blank, _np(labels) if hasattr(labels, "detach") else labels, hop_size_second)


def get_mae(gt, predict) -> float:
    """utils/alignment.py:190-199 (Python fp64, sequential order)."""
    error = 0.0
    cnt = 0
    for i in range(len(gt)):
        for j in range(len(gt[i])):
            error = error + abs(gt[i][j][0] - predict[i][j][0]) + abs(gt[i][j][1] - predict[i][j][1])
            cnt = cnt + 2.0
    return error / cnt
