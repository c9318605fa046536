"""Builds liblyricalign.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m lyricalignment_b200.build [--force]

The .so lands in lyricalignment_b200/_C/ (git-ignored, but it travels with the gpurun snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
SO_PATH = os.path.join(OUT_DIR, "liblyricalign.so")
SOURCES = ["la_emit.cu", "la_viterbi.cu", "la_logmel.cu", "la_head.cu", "la_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith("__")]
    paths.append(os.path.join(os.path.dirname(HERE), "include", "lyricalign.h"))
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= _newest_source_mtime():
        return SO_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(os.path.join(CSRC, "la_mel_table.inc")):      # committed; regenerate if missing
        subprocess.check_call([sys.executable, os.path.join(CSRC, "gen_mel_table.py")])
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", SO_PATH,
           *[os.path.join(CSRC, s) for s in SOURCES], "-lcuda"]
    subprocess.check_call(cmd)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
