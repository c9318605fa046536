"""Timeline of CTA 0's first tiles in K1 (LA_LOGMEL_DBG=8). Perf triage only."""
import ctypes, os, sys
os.environ["LA_LOGMEL_DBG"] = os.environ.get("LA_LOGMEL_DBG", "8")
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lyricalignment_b200 import _lib, audio as LA, synth
lib = _lib.load()
batch = synth.opencpop_shaped(400)
wave, off = synth.synthetic_waveforms(batch, device="cuda")
for _ in range(3):
    LA.log_mel_spectrogram_ragged(wave, off, batch.n_samples.astype(np.int32))
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * 8 * 32))()
lib.la_debug_logmel_trace.argtypes = [ctypes.c_void_p]
assert lib.la_debug_logmel_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(4, 8, 32)
t0 = t[t > 0].min()
names = ["mma", "xform", "epi", "prod"]
for tl in range(1, 6):
    print(f"--- tile {tl} (ns since first event)")
    m, x, e, p = t[0, tl] - t0, t[1, tl] - t0, t[2, tl] - t0, t[3, tl] - t0
    print(f" prod: raw_empty ok {p[0]}, raw issued {p[1]}, empty-wait ok ks0..4 {p[2:7].tolist()} ... ks24 {p[26]}")
    print(f" xform: raw_full ok {x[0]}, empty ok ks0..4 {x[1:6].tolist()} ... ks24 {x[25]}, done {x[26]}")
    print(f" mma: tmem_empty ok {m[0]}, full ok ks0..4 {m[1:6].tolist()} ... ks24 {m[25]}, committed {m[26]}")
    print(f" epi: tmem_full ok {e[0]}, w4 chunks done {e[2]}, w4 flushed {e[3]}, barrier passed {e[4]}, end {e[1]} | w8 chunks done {e[5]}, w8 flushed {e[6]}")
    print(f" xform first k-step: raw_full {x[0]}, edges staged {x[27]}, computed {x[28]}, stage free {x[1]}, stored {x[29]}, arrived {x[30]}")
    print(f" per-kstep mma deltas (ns): {np.diff(m[1:26]).tolist()}")
tl = 3
m, x, p = t[0, tl] - t0, t[1, tl] - t0, t[3, tl] - t0
print("ks : mma full-ok | xform(a_empty ok, even ks only) | producer b_empty ok")
for ks in range(25):
    print(f"{ks:2d} : {m[1+ks]:7d} | {x[1+ks] if ks % 2 == 0 else -1:7d} | {p[2+ks]:7d}")
