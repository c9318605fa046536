"""Per-config timings of the decode path (BASELINE.json configs 0/2/3: single 30 s clip, MIR-1k-shaped
songs through the CE decoder, 15 000 x 600 long-form trellis). Not the contract bench (bench.py);
prints one JSON line per config with K2 / K3 CUDA-event times."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lyricalignment_b200 import _lib, alignment as A, synth

dev = torch.device("cuda", 0)
lib = _lib.load()
V = synth.V_HEAD

def run(name, t_len, labels, ctc=True, reps=20):
    batch = synth.ClipBatch(np.asarray(t_len) * 0.02, np.asarray(t_len, np.int64) * 320, np.asarray(t_len, np.int32), labels)
    z = synth.planted_logits(batch, V, ctc=ctc, device=dev)
    l_len, cols = A._resolve_columns(labels, V - 2 if ctc else V - 1)
    plan = A.AlignPlan(A.MODE_CTC if ctc else A.MODE_CE, V, batch.t_len, l_len, cols, 0)
    ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)
    first = torch.empty(plan.total_labels, dtype=torch.int32, device=dev); last = torch.empty_like(first)
    score = torch.empty(plan.n_utt, dtype=torch.float64, device=dev); status = torch.empty(plan.n_utt, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    k2, k3 = [], []
    for i in range(reps + 3):
        flush.zero_()                                    # evict L2 between iterations
        e[0].record()
        _lib.check(lib.la_emit(plan.handle, z.data_ptr(), V, None, 0, ws.data_ptr(), st), "emit")
        e[1].record()
        _lib.check(lib.la_viterbi(plan.handle, ws.data_ptr(), first.data_ptr(), last.data_ptr(), score.data_ptr(), status.data_ptr(), st), "vit")
        e[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            k2.append(e[0].elapsed_time(e[1])); k3.append(e[1].elapsed_time(e[2]))
    assert int(status.max()) == 0
    audio = float(np.sum(t_len)) * 0.02
    k2m, k3m = float(np.median(k2)), float(np.median(k3))
    print(json.dumps({"config": name, "frames": int(np.sum(t_len)), "labels": int(l_len.sum()), "audio_s": audio,
                      "k2_ms": round(k2m, 4), "k3_ms": round(k3m, 4), "k2_gbs": round(4.0 * np.sum(t_len) * V / k2m / 1e6, 1),
                      "audio_s_per_s": round(audio / ((k2m + k3m) / 1e3), 1), "l2": "flushed between iterations"}))
    plan.close()

rng = np.random.default_rng(0)
lab = lambda L: rng.integers(2, 403, size=L).astype(np.int64)
run("c0/c1: single 30 s clip, T=1500, L=40, CTC", [1500], [lab(40)])
run("c0/c1: single 30 s clip, T=1500, L=40, CE", [1500], [lab(40)], ctc=False)
tl = rng.integers(1400, 5401, size=17); ll = [lab(int(rng.integers(48, 172))) for _ in tl]
run("c3: MIR-1k real shape (17 songs, T=1400-5400, L=48-171), CE decoder ('DTW' config)", tl, ll, ctc=False)
tl = rng.integers(200, 601, size=200); ll = [lab(int(np.clip(round(2.4 * t * 0.02), 1, t // 2))) for t in tl]
run("c3: BASELINE wording (200 clips of 4-12 s), CE decoder", tl, ll, ctc=False)
run("c4: long-form 15000 x 600, CTC", [15000], [lab(600)], reps=10)
