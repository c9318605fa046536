import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle
from oracle.logmel import log_mel_spectrogram_torch_f32
from lyricalignment_b200 import audio as LA
from test_gpu_logmel import _signal
for kind in ("survey", "noise", "tone"):
    rng = np.random.default_rng(5)
    a = _signal(rng, 16000 * 300, kind)
    want = oracle.log_mel_spectrogram(a)
    ours = LA.log_mel_spectrogram(a).cpu().numpy()
    ref32 = log_mel_spectrogram_torch_f32(a).numpy()
    for name, x in (("ours", ours), ("torch_f32", ref32)):
        e = 4 * np.abs(x - want)
        print(f"{kind:7s} {name:9s} log10-domain err: max {e.max():.2e} p99.99 {np.quantile(e, 0.9999):.2e} p99.9 {np.quantile(e, 0.999):.2e} median {np.median(e):.2e}  cells>1e-4: {(e > 1e-4).sum()} of {e.size}")
