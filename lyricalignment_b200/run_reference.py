"""Runs one of the reference's UNMODIFIED entry scripts on the B200 decode path:

    cd /path/to/LyricAlignment
    python -m lyricalignment_b200.run_reference inference_alignment.py -f test.json --model-dir ... --use-ctc-loss
    python -m lyricalignment_b200.run_reference inference_alignment_nogt.py -f songs.json --model-dir ...

Same CLI, same JSON inputs, same printed outputs. Before the script is executed the names it
imports are re-pointed:

  utils.alignment.perform_viterbi / perform_viterbi_ctc / get_mae   -> lyricalignment_b200.alignment
  whisper.audio.log_mel_spectrogram / pad_or_trim                   -> lyricalignment_b200.audio

(`from utils.alignment import ...` in inference_alignment.py:22 and `from whisper.audio import ...`
in module/align_model.py:9 then bind the CUDA versions.) The scripts' `align_logits.cpu()` keeps
working -- host logits are streamed back through the library's double-buffered host path (PCIe-bound);
deleting that one line (inference_alignment.py:161, inference_alignment_nogt.py:156) keeps the logits on
the GPU and is the only edit worth making.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types


def install(reference_root: str | None = None) -> dict:
    """Patches the modules in sys.modules; returns {patched name: original object}."""
    from . import alignment as la_align
    from . import audio as la_audio
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    originals = {}
    if "pypinyin" not in sys.modules:                     # imported but unused by utils/alignment.py:2
        try:
            importlib.import_module("pypinyin")
        except ImportError:
            stub = types.ModuleType("pypinyin")
            stub.lazy_pinyin = lambda *a, **k: []
            stub.Style = type("Style", (), {})
            sys.modules["pypinyin"] = stub
    ua = importlib.import_module("utils.alignment")
    for name in ("perform_viterbi", "perform_viterbi_ctc", "get_mae"):
        originals[f"utils.alignment.{name}"] = getattr(ua, name, None)
        setattr(ua, name, getattr(la_align, name))
    try:
        wa = importlib.import_module("whisper.audio")
    except ImportError:
        wa = None
    if wa is not None:
        for name in ("log_mel_spectrogram", "pad_or_trim"):
            originals[f"whisper.audio.{name}"] = getattr(wa, name, None)
            setattr(wa, name, getattr(la_audio, name))
    return originals


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    script = argv[0]
    root = os.path.dirname(os.path.abspath(script)) or os.getcwd()
    install(root)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
