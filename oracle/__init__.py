"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference alignment decode path (navi0105/LyricAlignment,
``utils/alignment.py`` and the un-vendored ``whisper.audio.log_mel_spectrogram``).
It exists to CHECK the CUDA path. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; nothing under
``lyricalignment_b200/`` does (tests/test_boundary.py greps for that).

Parity status
  * decode half (emission math, Viterbi DP, backtrace, on/offset, MAE): PINNED against
    the reference itself -- ``tests/golden/decode_*.npz`` were produced by importing
    ``/root/reference/utils/alignment.py`` unmodified (``tests/golden/make_golden.py``).
  * log-mel front end: the arithmetic lives in the third-party ``openai-whisper`` package
    (unpinned in the reference's requirements.txt:6, absent from /root/reference and from
    this image) -> restated from its published algorithm and cross-checked against
    ``transformers``' WhisperFeatureExtractor; **parity unpinned** w.r.t. the reference
    repo itself (it holds no test or fixture for it).
"""
from .viterbi import (align_one, viterbi_core, perform_viterbi, perform_viterbi_ctc, get_mae,
                      build as build_c_oracle)
from .emission import emission_ctc, emission_ce, emission_ctc_f64, emission_ce_f64
from .logmel import log_mel_spectrogram, mel_filterbank, decode_frames

__all__ = [
    "align_one", "viterbi_core", "perform_viterbi", "perform_viterbi_ctc", "get_mae",
    "build_c_oracle", "emission_ctc", "emission_ce", "emission_ctc_f64", "emission_ce_f64",
    "log_mel_spectrogram", "mel_filterbank", "decode_frames",
]
