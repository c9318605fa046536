import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _group(npz):
    cases = {}
    for key in npz.files:
        name, field = key.rsplit("/", 1)
        cases.setdefault(name, {})[field] = npz[key]
    return cases


@pytest.fixture(scope="session")
def decode_cases():
    return _group(np.load(os.path.join(GOLDEN, "decode_cases.npz")))


@pytest.fixture(scope="session")
def core_cases():
    return _group(np.load(os.path.join(GOLDEN, "core_cases.npz")))


@pytest.fixture(scope="session")
def logmel_cases():
    return _group(np.load(os.path.join(GOLDEN, "logmel_hf.npz")))


def split_onoff(onoff, lens):
    out, p = [], 0
    for n in lens:
        out.append([[float(a), float(b)] for a, b in onoff[p:p + int(n)]])
        p += int(n)
    return out
