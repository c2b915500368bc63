import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """libhp3d.so + host shim, compiled in-tree (nvcc cross-compiles without a GPU)."""
    from hierarchicalprobabilistic3dhuman_b200 import _build
    _build.build()
    return _build


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(a, b):
    """SURVEY.md §8d parity metric: max|a-b| / max|b|."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def reference_config():
    from types import SimpleNamespace as NS
    return NS(MODEL=NS(NUM_IN_CHANNELS=18, NUM_RESNET_LAYERS=18, EMBED_DIM=256, DELTA_I=True, DELTA_I_WEIGHT=1.0,
                       NUM_SMPL_BETAS=10))
