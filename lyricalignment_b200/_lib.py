"""ctypes binding of include/lyricalign.h. There is no CPU fallback: if the library is missing
or no CUDA device is visible, every compute call raises."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "_C", "liblyricalign.so")

LA_OK = 0
MODE_CTC, MODE_CE, MODE_LOGP = 0, 1, 2
UTT_OK, UTT_EMPTY, UTT_INFEASIBLE = 0, 1, 2
MAX_LABELS = 8191

# every symbol include/lyricalign.h declares (tests/test_boundary.py checks the header against this)
_c_int, _i32p, _i64, _vp, _sz = ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t
SIGNATURES = {
    "la_version": (ctypes.c_char_p, []),
    "la_last_error": (ctypes.c_char_p, []),
    "la_shutdown": (None, []),
    "la_device_count": (_c_int, []),
    "la_plan_create": (_c_int, [ctypes.POINTER(_vp), _c_int, _c_int, _c_int, _vp, _vp, _vp, _c_int]),
    "la_plan_destroy": (None, [_vp]),
    "la_plan_workspace_bytes": (_sz, [_vp]),
    "la_plan_total_frames": (_i64, [_vp]),
    "la_plan_total_labels": (_i64, [_vp]),
    "la_plan_num_launches": (_c_int, [_vp]),
    "la_plan_utt_layout": (_c_int, [_vp, _c_int, ctypes.POINTER(_i64), ctypes.POINTER(ctypes.c_int32),
                                    ctypes.POINTER(_i64), ctypes.POINTER(ctypes.c_int32)]),
    "la_plan_utt_bp_layout": (_c_int, [_vp, _c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "la_emit": (_c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "la_viterbi": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "la_viterbi_debug": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "la_align": (_c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "la_align_host": (_c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz]),
    "la_head_packed_weight_bytes": (_sz, [_c_int, _c_int]),
    "la_head_pack_weights": (_c_int, [_vp, _i64, _c_int, _c_int, _vp, _vp]),
    "la_head_workspace_bytes": (_sz, [_vp, _c_int]),
    "la_head_emit": (_c_int, [_vp, _vp, _i64, _c_int, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "la_head_align": (_c_int, [_vp, _vp, _i64, _c_int, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "la_logmel_workspace_bytes": (_sz, [_c_int, _i64]),
    "la_logmel": (_c_int, [_vp, _c_int, _i64, _i64, _vp, _i64, _vp, _vp]),
    "la_logmel_ragged": (_c_int, [_vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class LyricAlignError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads liblyricalign.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build the CUDA library first "
                "(python -m lyricalignment_b200.build, or __graft_entry__.build()). "
                "lyricalignment_b200 has no CPU fallback.")
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != LA_OK:
        msg = load().la_last_error().decode("utf-8", "replace")
        raise LyricAlignError(f"{what} failed (status {rc}): {msg}")
