"""Share of wall time: product kernels (K1, K2+K3) vs the stock encoder + head that sit between them
(SURVEY.md 8d). Random-init Whisper-tiny / -medium stand-ins, fp32, one 30 s clip and a batch."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lyricalignment_b200 as la
from lyricalignment_b200 import audio as LA, synth
from lyricalignment_b200.pipeline import AlignPipeline

def ev(): return torch.cuda.Event(enable_timing=True)

def run(size, n_clips, seconds):
    pipe = AlignPipeline(size, vocab=synth.V_HEAD)
    rng = np.random.default_rng(0)
    audios = [(0.1 * rng.standard_normal(int(16000 * seconds))).astype(np.float32) for _ in range(n_clips)]
    labels = [[int(x) for x in rng.integers(2, 403, size=max(1, int(2.4 * seconds)))] for _ in range(n_clips)]
    res = {}
    for it in range(4):
        t = [ev() for _ in range(4)]
        wav = torch.from_numpy(np.stack(audios)).cuda()
        t[0].record()
        mel = LA.log_mel_spectrogram(wav)
        t[1].record()
        F_ = mel.shape[-1]
        with torch.no_grad():
            emb = pipe.embed_audio(LA.pad_or_trim(mel, LA.N_FRAMES))[:, :LA.decode_frames(F_), :]
            logits = pipe.head(emb)
        t[2].record()
        out = la.perform_viterbi_ctc(logits, labels)
        t[3].record()
        torch.cuda.synchronize()
        res = {"k1_ms": t[0].elapsed_time(t[1]), "encoder_head_ms": t[1].elapsed_time(t[2]), "k2_k3_and_readback_ms": t[2].elapsed_time(t[3])}
    tot = sum(res.values())
    print(json.dumps({"model": f"whisper-{size} stand-in (random init, fp32) + GRU head, V=21129", "clips": n_clips,
                      "seconds_each": seconds, **{k: round(v, 3) for k, v in res.items()},
                      "product_share": round((res["k1_ms"] + res["k2_k3_and_readback_ms"]) / tot, 4),
                      "audio_s_per_s_end_to_end": round(n_clips * seconds / (tot / 1e3), 1)}))

run("tiny", 1, 30.0)
run("tiny", 16, 10.0)
run("medium", 1, 30.0)
run("medium", 16, 10.0)
