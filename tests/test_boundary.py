"""CPU: the C-ABI library loads and exports every symbol include/lyricalign.h declares, the
Python mirror keeps the reference's names/signatures, and the product never touches oracle/."""
import ctypes
import inspect
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "lyricalign.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(la_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from lyricalignment_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lyricalign.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms
    assert b"sm_100a" in lib.la_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import lyricalignment_b200 as la
    from lyricalignment_b200._lib import LyricAlignError
    with pytest.raises(LyricAlignError):
        la.perform_viterbi_ctc(np.zeros((1, 4, 12), np.float32), [[3]])
    # and the C ABI itself refuses rather than computing on the host
    from lyricalignment_b200 import _lib
    h = ctypes.c_void_p()
    t = (ctypes.c_int32 * 1)(4); l = (ctypes.c_int32 * 1)(1); c = (ctypes.c_int32 * 1)(3)
    rc = _lib.load().la_plan_create(ctypes.byref(h), 0, 1, 12, t, l, c, 0)
    assert rc == -2 and _lib.load().la_last_error()


def test_product_never_imports_oracle_or_reference():
    pkg = os.path.join(ROOT, "lyricalignment_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f
                assert "oracle/" not in src and "liboracle" not in src, f


def test_python_mirror_keeps_reference_signatures():
    import lyricalignment_b200.alignment as a
    for name in ("perform_viterbi_ctc", "perform_viterbi"):
        sig = inspect.signature(getattr(a, name))
        assert list(sig.parameters) == ["prediction", "labels", "hop_size_second"]
        assert sig.parameters["hop_size_second"].default == 0.02
    assert list(inspect.signature(a.get_mae).parameters) == ["gt", "predict"]
    assert list(inspect.signature(a.run_viterbi_core).parameters) == [
        "dp_matrix", "backtrace_dp_matrix", "cur_log_prediction", "cur_log_silence_prediction", "cur_label"]


def test_get_mae_matches_golden():
    import json
    from conftest import GOLDEN
    import lyricalignment_b200 as la
    g = json.load(open(os.path.join(GOLDEN, "mae_cases.json")))
    assert la.get_mae([[[0, .5], [.5, 1]]], [[[.02, .48], [.5, 1.02]]]) == 0.01500000000000001
    for gt, pr, v in zip(g["gt"], g["predict"], g["mae"]):
        assert la.get_mae(gt, pr) == v


def test_label_resolution_follows_numpy_indexing():
    import numpy as np
    from lyricalignment_b200.alignment import _label_rows, _resolve_columns
    rows = _label_rows(np.array([[3, 4, -100], [7, -100, -100]]))
    lens, cols = _resolve_columns(rows, 10)
    assert lens.tolist() == [2, 1] and cols.tolist() == [3, 4, 7]
    # label 0 -> emission column -1 -> wraps to the last column (numpy semantics)
    assert _resolve_columns([np.array([0])], 10)[1].tolist() == [10]
    with pytest.raises(IndexError):
        _resolve_columns([np.array([11])], 10)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: it times the reference's CPU path (the unmodified
    reference where its tree is mounted, else the oracle port) and prints ONE JSON line."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-s/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
