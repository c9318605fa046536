import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, oracle
from lyricalignment_b200 import audio as LA
from test_gpu_logmel import _signal
for n, kind in [(3333, "silence_tail"), (3333, "noise"), (641, "silence_tail"), (16000, "silence_tail")]:
    rng = np.random.default_rng(n % 97)
    a = _signal(rng, n, kind)
    for rep in range(3):
        got = LA.log_mel_spectrogram(a).cpu().numpy(); want = oracle.log_mel_spectrogram(a)
        e = np.abs(got - want); m, f = np.unravel_index(e.argmax(), e.shape)
        print(n, kind, rep, "max err", e.max(), "at mel", m, "frame", f, "got", got[m, f], "want", want[m, f], "n bad", (e > 1e-4).sum(), "bad mels", sorted(set(np.nonzero(e > 1e-4)[0].tolist()))[:10])
