"""oracle/ref_shim.py -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference decode module ``/root/reference/utils/alignment.py`` so it
can be used (a) to generate the golden fixtures under tests/golden/ and (b) as the
``"kind": "reference"`` CPU arm when the reference tree is present. The reference imports
``pypinyin`` (utils/alignment.py:2) without using it; the package is not installed, so a
stub module is registered first. /root/reference does not exist on the GPU box: callers must
check ``available()`` and fall back to the oracle port.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LA_REFERENCE_ROOT", "/root/reference")
_mod = None


def available() -> bool:
    if not os.path.exists(os.path.join(REFERENCE_ROOT, "utils", "alignment.py")):
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
            raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
        if "pypinyin" not in sys.modules:
            stub = types.ModuleType("pypinyin")
            stub.lazy_pinyin = lambda *a, **k: []
            stub.Style = type("Style", (), {})
            sys.modules["pypinyin"] = stub
        spec = importlib.util.spec_from_file_location(
            "_la_reference_alignment", os.path.join(REFERENCE_ROOT, "utils", "alignment.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mod = mod
    return _mod
