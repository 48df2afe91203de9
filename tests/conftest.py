import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

CFG_6M = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
CFG_94M = dict(dimension=3, input_nc=1, output_nc=32, num_downs=5, ngf=32,
               norm="instance", pooling="Avg", interp="trilinear", norm_eps=1e-2)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rand_input(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def state_6m():
    z = golden("anatomix_6m_state.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def min_cosine(a, b):
    """Smallest per-voxel cosine similarity across channels ([N,C,...] tensors)."""
    a, b = a.double(), b.double()
    num = (a * b).sum(1)
    den = a.norm(dim=1) * b.norm(dim=1)
    return (num / den.clamp_min(1e-30)).min().item()
