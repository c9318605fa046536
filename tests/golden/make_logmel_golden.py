"""Generates tests/golden/logmel_hf_long.npz: the log-mel oracle's second and third independent anchors.

openai-whisper (where the reference's front-end arithmetic lives, module/align_model.py:9,84) is not
installed here and is unpinned in the reference's requirements.txt, so the oracle (oracle/logmel.py, fp64)
is a restatement. This script pins it against the two implementations of the same published algorithm that
ARE in this image, on the shapes the benchmark configs use:
  * transformers' WhisperFeatureExtractor._np_extract_fbank_features  (numpy STFT, fp64 inside)
  * transformers' WhisperFeatureExtractor._torch_extract_fbank_features (torch.stft fp32 -- the very
    formulation whisper.audio.log_mel_spectrogram uses)
for a 30 s clip (BASELINE config 1) and a 5-minute song (config 4). To keep the fixture small only every
STRIDE-th frame is stored; the waveform is regenerated from its seed by the test.

    python tests/golden/make_logmel_golden.py
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STRIDE = {"c1_30s": 5, "c4_300s": 50}


def signal(name):
    seconds, seed = {"c1_30s": (30, 11), "c4_300s": (300, 12)}[name]
    rng = np.random.default_rng(seed)
    n = 16000 * seconds
    t = np.arange(n) / 16000.0
    a = 0.1 * rng.standard_normal(n)
    for h in range(1, 11):
        a += (0.3 / h) * np.sin(2 * np.pi * 220 * h * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t))
    a[int(0.9 * n):] = 0.0
    return a.astype(np.float32)


def main():
    import torch
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor()
    out = {}
    for name, stride in STRIDE.items():
        a = signal(name)
        f_np = fe._np_extract_fbank_features(a[None], "cpu")[0]            # one sample: per-sample max == global max
        f_t = fe._torch_extract_fbank_features(a[None], "cpu")[0]
        assert f_np.shape == f_t.shape == (80, len(a) // 160)
        out[f"{name}/hf_numpy"] = np.ascontiguousarray(f_np[:, ::stride]).astype(np.float32)
        out[f"{name}/hf_torch_f32"] = np.ascontiguousarray(f_t[:, ::stride]).astype(np.float32)
        out[f"{name}/stride"] = np.int64(stride)
        print(name, f_np.shape, "np-vs-torch max", float(np.abs(f_np - f_t).max()))
    np.savez_compressed(os.path.join(HERE, "logmel_hf_long.npz"), **out)


if __name__ == "__main__":
    main()
