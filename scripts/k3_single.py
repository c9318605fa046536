"""K3 alone on a small-V plan (for ncu source-level captures and quick timings):
    python scripts/k3_single.py T L [reps] [n_clips]       n identical clips
    python scripts/k3_single.py opencpop N [reps]          the bench's Opencpop-shaped batch (N clips of 5-15 s)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lyricalignment_b200 import _lib, alignment as A, synth

dev = torch.device("cuda", 0)
lib = _lib.load()
rng = np.random.default_rng(0)
V = 512
if sys.argv[1] == "opencpop":
    n = int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    batch = synth.opencpop_shaped(n)
    labels, t_len = batch.labels, batch.t_len
    T, L = int(t_len.sum()), sum(len(x) for x in labels)
else:
    T, L = int(sys.argv[1]), int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    labels = [rng.integers(2, 403, size=L).astype(np.int64) for _ in range(n)]
    t_len = np.full(n, T, np.int32)
    batch = synth.ClipBatch(t_len * 0.02, t_len.astype(np.int64) * 320, t_len, labels)
z = synth.planted_logits(batch, V, ctc=True, device=dev)
l_len, cols = A._resolve_columns(labels, V - 2)
plan = A.AlignPlan(A.MODE_CTC, V, t_len, l_len, cols, 0)
ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
first = torch.empty(plan.total_labels, dtype=torch.int32, device=dev); last = torch.empty_like(first)
score = torch.empty(n, dtype=torch.float64, device=dev); status = torch.empty(n, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.la_emit(plan.handle, z.data_ptr(), V, None, 0, ws.data_ptr(), st), "emit")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(reps):
    e0.record()
    _lib.check(lib.la_viterbi(plan.handle, ws.data_ptr(), first.data_ptr(), last.data_ptr(), score.data_ptr(), status.data_ptr(), st), "vit")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
assert int(status.max()) == 0
print(f"T={T} L={L} n={n}: K3 {np.median(ts) * 1e3:.1f} us (median of {reps})")
