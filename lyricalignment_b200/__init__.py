"""B200-native alignment decode path for navi0105/LyricAlignment (drop-in for utils/alignment.py
and whisper.audio.log_mel_spectrogram). Hand-written sm_100a CUDA behind a C ABI; no CPU fallback."""
from .alignment import (AlignJob, AlignPlan, AlignResult, align, align_clips, align_clips_async, get_mae,
                        onoff_seconds, perform_viterbi, perform_viterbi_ctc, run_viterbi_core)

__all__ = ["AlignJob", "AlignPlan", "AlignResult", "align", "align_clips", "align_clips_async", "get_mae",
           "onoff_seconds", "perform_viterbi", "perform_viterbi_ctc", "run_viterbi_core"]
__version__ = "0.1.0"
