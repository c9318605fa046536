"""N1 timing: fused head (pack + tcgen05 GEMM/LSE + fp32 gather) + K3 on the 2 000-clip workload's hidden states,
next to what it replaces: the stock fp32 Linear (cuBLAS, TF32 off = PyTorch's default, and TF32 on) writing the
[sum T][V] logits + K2 reading them. Prints one JSON line."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lyricalignment_b200 import _lib, alignment as A, synth
from lyricalignment_b200.head import FusedHead

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda", 0)
torch.manual_seed(114514)
batch = synth.opencpop_shaped(clips)
V, D = synth.V_HEAD, 768
T = int(batch.t_len.sum())
fc = torch.nn.Linear(D, V).to(dev)
X = torch.nn.functional.mish(torch.randn(T, D, device=dev))          # what the head's Linear sees (align_model.py:38)
head = FusedHead(fc.weight, fc.bias)
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

t_emit = [ev(), ev()]
def fused():
    job = head.align_clips_async(X, batch.t_len, batch.labels, timing=t_emit)
    r = job.result(); job.close(); return r
ms_fused = timed(fused)
ms_head_emit = t_emit[0].elapsed_time(t_emit[1])
# what it replaces, on a row subset that fits (the full logits are 84.5 GB): stock Linear + K2
rows = min(T, 200_000)
sub_t = []
acc = 0
for t in batch.t_len:
    if acc + int(t) > rows: break
    sub_t.append(int(t)); acc += int(t)
rows = acc; nsub = len(sub_t)
Xs = X[:rows]
out = {}
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    logits = None
    def stock():
        global logits
        logits = torch.addmm(fc.bias, Xs, fc.weight.t())
    ms_fc = timed(stock)
    ms_k2k3 = timed(lambda: A.align_clips(logits, np.array(sub_t, np.int32), batch.labels[:nsub]))
    out["tf32_on" if tf32 else "fp32"] = {"fc_ms_scaled_to_full": round(ms_fc * T / rows, 2), "k2_k3_ms_scaled": round(ms_k2k3 * T / rows, 2)}
torch.backends.cuda.matmul.allow_tf32 = False
flops = 2.0 * T * V * D
print(json.dumps({"clips": clips, "frames": T, "V": V, "D": D,
                  "fused_total_ms": round(ms_fused, 2), "fused_head_emit_ms": round(ms_head_emit, 2),
                  "fused_tflops_useful": round(flops / (ms_head_emit / 1e3) / 1e12, 1),
                  "fused_tflops_issued_fp16": round(3 * flops / (ms_head_emit / 1e3) / 1e12, 1),
                  "audio_s_per_s": round(batch.audio_seconds / (ms_fused / 1e3), 1),
                  "stock_linear_then_k2": out, "subset_rows_for_stock": rows}))
