"""K3 phase clocks (LA_VIT_TRACE=1): cycles spent in the DP and in the backtrace for a single 30 s clip (T=1500,
L=40), a single-warp clip (T=600, L=24) and the long-form trellis (15000 x 600); more shapes: T:L pairs as arguments.
Perf triage only."""
import ctypes, os, sys
os.environ["LA_VIT_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lyricalignment_b200 import _lib, alignment as A, synth
lib = _lib.load()
lib.la_debug_viterbi_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
V = 512
shapes = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(1500, 40), (600, 24), (15000, 600)]
for T, L in shapes:
    lab = rng.integers(2, 403, size=L).astype(np.int64)
    batch = synth.ClipBatch(np.array([T * 0.02]), np.array([T * 320]), np.array([T], np.int32), [lab])
    z = synth.planted_logits(batch, V, ctc=True, device=dev)
    for _ in range(3):
        A.align_clips(z, batch.t_len, [lab])
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 4)()
    assert lib.la_debug_viterbi_trace(buf) == 0
    c0, c1, c2, t = [int(x) for x in buf]
    print(f"T={T} L={L}: DP {c1 - c0} cycles = {(c1 - c0) / t:.1f} per frame; backtrace {c2 - c1} cycles = {(c2 - c1) / t:.1f} per frame")
