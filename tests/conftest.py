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


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def nao():
    g = load_golden("nao.npz")
    cpl = g["complete_pc_list"]
    ci = int(g["cano_idx"])
    cano = cpl[ci]
    pc_list = np.concatenate([cpl[:ci], cpl[ci + 1:]])
    return g, cano, pc_list


def synthetic_sequence(T, N, P, seed=2, M=None):
    """Small articulated synthetic sequence (same generator family as bench.py, see reart_b200/synth.py)."""
    from reart_b200.synth import make_sequence
    return make_sequence(T=T, N=N, P=P, seed=seed, M=M)
