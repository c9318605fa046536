import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from oracle.logmel import log_mel_spectrogram_torch_f32
from lyricalignment_b200 import audio as LA
rng = np.random.default_rng(2)
def sig(n, kind, amp=1.0):
    t = np.arange(n) / 16000.0
    if kind == "tone": a = 0.8 * np.sin(2 * np.pi * 440.0 * t)
    elif kind == "tone+noise": a = 0.8 * np.sin(2 * np.pi * 440.0 * t) + 1e-4 * rng.standard_normal(n)
    elif kind == "noise": a = 0.1 * rng.standard_normal(n)
    else:
        a = 0.1 * rng.standard_normal(n)
        for h in range(1, 11): a += (0.3 / h) * np.sin(2 * np.pi * 220 * h * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t))
        a[int(0.9 * n):] = 0.0
    return (amp * a).astype(np.float32)
for kind in ["noise", "survey", "tone", "tone+noise"]:
    for n in [16000 * 5 + 37, 16000 * 15 - 1, 480000]:
        for amp in [1.0, 0.07]:
            a = sig(n, kind, amp)
            want = oracle.log_mel_spectrogram(a)
            ours = LA.log_mel_spectrogram(a).cpu().numpy()
            ref32 = log_mel_spectrogram_torch_f32(a).numpy()
            e1, e2 = np.abs(ours - want), np.abs(ref32 - want)
            print(f"{kind:11s} n={n:7d} amp={amp:4.2f} ours-vs-f64 max {4*e1.max():.2e} p99.9 {4*np.quantile(e1,0.999):.2e} | torchf32-vs-f64 max {4*e2.max():.2e} p99.9 {4*np.quantile(e2,0.999):.2e} | ours-vs-torchf32 max {4*np.abs(ours-ref32).max():.2e}  (log10 units)")
print("---- where is the worst error? ----")
for kind, n, amp in [("survey", 80037, 0.07), ("survey", 239999, 1.0), ("noise", 480000, 1.0)]:
    rng = np.random.default_rng(2)
    a = sig(n, kind, amp)
    want = oracle.log_mel_spectrogram(a); ours = LA.log_mel_spectrogram(a).cpu().numpy()
    ref32 = log_mel_spectrogram_torch_f32(a).numpy()
    e = np.abs(ours - want)
    idx = np.dstack(np.unravel_index(np.argsort(-e.ravel())[:6], e.shape))[0]
    for m, f in idx:
        print(kind, n, amp, "mel", m, "frame", f, "/", want.shape[1], "want*4-4 (log10)", 4*want[m, f]-4, "max log10", 4*want.max()-4, "err_ours", 4*e[m, f], "err_ref32", 4*abs(ref32[m,f]-want[m,f]))
