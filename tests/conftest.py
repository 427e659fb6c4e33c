import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def sd_g():
    from glare_b200 import synth
    return synth.synth_state_dict("netG", 0)


@pytest.fixture(scope="session")
def sd_v():
    from glare_b200 import synth
    return synth.synth_state_dict("vqgan", 0)


@pytest.fixture(scope="session")
def glare_lib():
    """Build (if needed) and load the C-ABI library; GPU tests call the product only through it."""
    from glare_b200 import build, _lib
    build.build()
    return _lib.lib()
