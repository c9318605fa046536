"""oracle/emission.py -- TEST INFRASTRUCTURE ONLY (CPU checker, never the product path).

numpy restatement of the reference's emission math:
  * CTC flavour  -- utils/alignment.py:123-134 (``perform_viterbi_ctc``)
  * CE flavour   -- utils/alignment.py:14-20   (``perform_viterbi``; the "DTW" config)

The fp32 versions mirror the reference's operation ORDER in float32 (max, exp, sum, log,
``(z - max) - log_sum``, naive ``1/(1+exp(-z))`` sigmoid, ``log(1 - s)``, add, then clip at
-1000). numpy's libm and torch's Sleef differ by an ulp here and there, so these agree with
the reference to ~1e-6 absolute outside the sigmoid-saturation band (checked against the
reference's own output in tests/golden/decode_*.npz by tests/test_oracle.py); the fp64
versions are the tolerance anchor for the CUDA kernel.
"""
from __future__ import annotations

import numpy as np

CLIP = -1000.0


def _log_softmax(z: np.ndarray, dtype) -> np.ndarray:
    z = z.astype(dtype, copy=False)
    m = z.max(axis=-1, keepdims=True)
    d = z - m
    s = np.exp(d).sum(axis=-1, keepdims=True, dtype=dtype)
    return (d - np.log(s)).astype(dtype, copy=False)


def _emission_ctc(pred: np.ndarray, dtype):
    """pred [..., T, V] -> (emit [..., T, V-2], blank [..., T, 1]); utils/alignment.py:123-134."""
    pred = np.asarray(pred)
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        lp = _log_softmax(pred[..., 1:-1], dtype)                    # :123
        zs = pred[..., -1:].astype(dtype, copy=False)
        one = dtype(1.0)
        s = one / (one + np.exp(-zs))                                 # :125 (F.sigmoid)
        voiced = one - s                                              # :126
        log_s = np.log(s)                                             # :128
        log_v = np.log(voiced)                                        # :129
        emit = np.maximum(lp + log_v, dtype(CLIP))                    # :131-132
        blank = np.maximum(log_s, dtype(CLIP))                        # :134
    return emit, blank


def _emission_ce(pred: np.ndarray, dtype):
    """pred [..., T, V] -> (emit [..., T, V-1], blank [..., T, 1]); utils/alignment.py:14-20."""
    lp = _log_softmax(np.asarray(pred), dtype)                        # :14
    blank = np.maximum(lp[..., 0:1], dtype(CLIP))                     # :16,20
    emit = np.maximum(lp, dtype(CLIP))[..., 1:]                       # :18
    return emit, blank


def emission_ctc(pred):
    return _emission_ctc(pred, np.float32)


def emission_ce(pred):
    return _emission_ce(pred, np.float32)


def emission_ctc_f64(pred):
    return _emission_ctc(pred, np.float64)


def emission_ce_f64(pred):
    return _emission_ce(pred, np.float64)
