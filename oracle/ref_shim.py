"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference decode module ``/root/reference/utils/alignment.py`` so it
can be used (a) to generate the golden fixtures under tests/golden/ and (b) as the
``"kind": "reference"`` CPU arm when the reference tree is present. The reference imports
``pypinyin`` (utils/alignment.py:2) without using it; the package is not installed, so a
stub module is registered first.

/root/reference does not exist on the GPU box. ``vendor()`` (called by ``__graft_entry__.build()`` in
the dev container, where the tree is mounted) therefore copies that ONE file, byte for byte, to
``oracle/_ref/utils/alignment.py``. ``oracle/_ref/`` is git-ignored (reference sources never enter the
history) but travels with the gpurun snapshot like the built .so files, so ``bench.py``'s CPU arm on the
GPU box times the unmodified reference (``"kind": "reference"``). Callers still check ``available()``
and fall back to the oracle port when neither location exists.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LA_REFERENCE_ROOT", "/root/reference")
VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mod = None


def _source() -> str | None:
    for root in (REFERENCE_ROOT, VENDORED_ROOT):
        p = os.path.join(root, "utils", "alignment.py")
        if os.path.exists(p):
            return p
    return None


def vendor() -> str | None:
    """Copies the reference's utils/alignment.py unmodified into oracle/_ref/ (git-ignored). Returns the
    destination, or None when the reference tree is not mounted."""
    import shutil
    src = os.path.join(REFERENCE_ROOT, "utils", "alignment.py")
    if not os.path.exists(src):
        return None
    dst = os.path.join(VENDORED_ROOT, "utils", "alignment.py")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    return dst


def available() -> bool:
    if _source() is None:
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def load():
    """Returns the reference's utils.alignment module (perform_viterbi, perform_viterbi_ctc,
    run_viterbi_core, get_mae)."""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError(f"reference decode module found neither under {REFERENCE_ROOT} nor {VENDORED_ROOT}")
        if "pypinyin" not in sys.modules:
            stub = types.ModuleType("pypinyin")
            stub.lazy_pinyin = lambda *a, **k: []
            stub.Style = type("Style", (), {})
            sys.modules["pypinyin"] = stub
        spec = importlib.util.spec_from_file_location(
            "_la_reference_alignment", _source())
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mod = mod
    return _mod
