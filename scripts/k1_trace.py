"""Timeline of CTA 0's first tiles in K1 (LA_LOGMEL_DBG=8). Perf triage only.
Events (la_logmel.cu `trace`): role 0 = cross-term MMA issuer: 0/29 TMEM free pass 0/1, 1+i operands of
k-step i ready (i = 0..25), 27/28 pass 0/1 committed; role 1 = transform warp 12: 0 raw_full, 1 scale known,
2+j its j-th k-step's TMEM stage free, 30 tile done; role 2 = epilogue warp 4: 0/2 tmem_full pass 0/1,
1/3 drained pass 0/1, 4 finish done; role 3 = producer: 0 raw_empty ok, 1 raw issued, 2+i basis stage free."""
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
buf = (ctypes.c_ulonglong * (4 * 8 * 64))()
lib.la_debug_logmel_trace.argtypes = [ctypes.c_void_p]
assert lib.la_debug_logmel_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(4, 8, 64)
t0 = t[t > 0].min()
for tl in range(1, 6):
    m, x, e, p = t[0, tl] - t0, t[1, tl] - t0, t[2, tl] - t0, t[3, tl] - t0
    print(f"--- tile {tl} (ns since first event); tile period {t[0, tl, 0] - t[0, tl - 1, 0]} ns")
    print(f" prod : raw_empty ok {p[0]}, raw issued {p[1]}, basis stage free k0..3 {p[2:6].tolist()} .. k25 {p[27]}")
    print(f" xform: raw_full ok {x[0]}, edges done {x[32]}, scan loop done {x[33]}, scanned {x[31]}, barrier passed {x[29]}, scale known {x[1]}, tile done {x[30]}")
    print(f"        own k-steps computed {x[16:23].tolist()}")
    print(f"        own stages free      {x[2:9].tolist()}")
    print(f" mma  : pass0 TMEM free {m[0]}, pass0 committed {m[27]} | pass1 TMEM free {m[29]}, pass1 committed {m[28]}")
    print(f" epi  : pass0 full {e[0]} drained {e[1]} | pass1 full {e[2]} drained {e[3]} finish done {e[4]}")
    print(f" issuer k-steps 4..7: [start, 4 MMAs issued, look-ahead done, 5th issued, committed] relative to k4 start:")
    for j in range(4):
        print(f"        k{4+j}: {(m[32+5*j:37+5*j] - m[32]).tolist()}")
