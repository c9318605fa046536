"""K1 bring-up check (GPU): max |ours - fp64 oracle| on two short clips, with the fp16 pair order as
designed and swapped (LA_LOGMEL_DBG=16). Prints which one is right. Bounded: a hang is killed by the caller's timeout."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from lyricalignment_b200 import audio as LA
rng = np.random.default_rng(0)
for n in (16000 * 3, 16000 * 7 + 123):
    t = np.arange(n) / 16000.0
    a = (0.1 * rng.standard_normal(n) + 0.3 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    want = oracle.log_mel_spectrogram(a)
    for dbg in ("0", "16"):
        os.environ["LA_LOGMEL_DBG"] = dbg
        got = LA.log_mel_spectrogram(a).cpu().numpy()
        e = 4 * np.abs(got - want)
        print(f"n={n} LA_LOGMEL_DBG={dbg}: max err {e.max():.3e} median {np.median(e):.3e} (log10 units) nan={np.isnan(got).sum()}", flush=True)
os.environ["LA_LOGMEL_DBG"] = "0"
