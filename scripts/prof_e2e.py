"""Where does a per-clip host-path call spend its time? (run on the GPU box)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lyricalignment_b200 as la
from lyricalignment_b200 import alignment as A, synth, _lib

dev = torch.device("cuda", 0)
batch = synth.opencpop_shaped(32)
V = synth.V_HEAD
logits = synth.planted_logits(batch, V, device=dev)
host = torch.empty(logits.shape, dtype=torch.float32).pin_memory(); host.copy_(logits); torch.cuda.synchronize()
offs = np.concatenate([[0], np.cumsum(batch.t_len)])
# raw H2D bandwidth
big = torch.empty(1 << 28, dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device=dev)
for _ in range(2): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s", 3 * big.numel() * 4 / (time.perf_counter() - t0) / 1e9)
del big, dbig
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
preds = [host[offs[i]:offs[i+1]].unsqueeze(0) for i in range(32)]
labs = [batch.labels[i].tolist() for i in range(32)]
print("perform_viterbi_ctc per clip ms", 1e3 * t(lambda: [la.perform_viterbi_ctc(preds[i], [labs[i]]) for i in range(32)]) / 32)
rows = [batch.labels[i] for i in range(32)]
def plans():
    for i in range(32):
        l_len, cols = A._resolve_columns([rows[i]], V - 2)
        p = A.AlignPlan(A.MODE_CTC, V, np.array([batch.t_len[i]], np.int32), l_len, cols, 0); p.close()
print("plan create+destroy ms", 1e3 * t(plans) / 32)
ps = []
for i in range(32):
    l_len, cols = A._resolve_columns([rows[i]], V - 2)
    ps.append(A.AlignPlan(A.MODE_CTC, V, np.array([batch.t_len[i]], np.int32), l_len, cols, 0))
print("la_align_host only ms", 1e3 * t(lambda: [A._run_host(ps[i], preds[i][0]) for i in range(32)]) / 32)
for sb in (4 << 20, 16 << 20, 64 << 20, 256 << 20):
    print("  staging", sb >> 20, "MiB ms", 1e3 * t(lambda: [A._run_host(ps[i], preds[i][0], sb) for i in range(32)]) / 32)
print("mean clip MB", float(np.mean(batch.t_len)) * V * 4 / 1e6)
# batched host call
l_len, cols = A._resolve_columns(rows, V - 2)
pb = A.AlignPlan(A.MODE_CTC, V, batch.t_len, l_len, cols, 0)
for sb in (16 << 20, 64 << 20):
    dt = t(lambda: A._run_host(pb, host, sb)); print("batched 32 clips staging", sb >> 20, "ms", dt * 1e3, "GB/s", host.numel() * 4 / dt / 1e9)
