"""Where does the host time of one FusedHead.align_clips_async() call go? (perf triage)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lyricalignment_b200 import _lib, alignment as A, synth, head as H
dev = torch.device("cuda", 0)
batch = synth.opencpop_shaped(2000)
V, D = synth.V_HEAD, 768
T = int(batch.t_len.sum())
fc = torch.nn.Linear(D, V).to(dev)
X = torch.nn.functional.mish(torch.randn(T, D, device=dev))
head = H.FusedHead(fc.weight, fc.bias)
lib = _lib.load()
for rep in range(3):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    lens, flat = A._flatten_labels(batch.labels); l_len, cols = A._resolve_columns((lens, flat), V - 2); t.append(time.perf_counter())
    plan = A.AlignPlan(A.MODE_CTC, V, batch.t_len, l_len, cols, 0); t.append(time.perf_counter())
    ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
    hws = torch.empty(int(lib.la_head_workspace_bytes(plan.handle, D)), dtype=torch.uint8, device=dev); t.append(time.perf_counter())
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.la_head_emit(plan.handle, X.data_ptr(), X.stride(0), D, head.weight.data_ptr(), head.weight.stride(0),
                                head.bias.data_ptr(), head.packed.data_ptr(), hws.data_ptr(), ws.data_ptr(), st), "emit"); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    plan.close(); del ws, hws; t.append(time.perf_counter())
    print("rep", rep, "labels %.1f ms, plan %.1f, alloc %.1f, enqueue %.1f, gpu wait %.1f, close %.1f" % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    job = head.align_clips_async(X, batch.t_len, batch.labels); t1 = time.perf_counter()
    r = job.result(); t2 = time.perf_counter(); job.close(); t3 = time.perf_counter()
    print("job rep", rep, "async %.1f ms, result %.1f, close %.1f" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
