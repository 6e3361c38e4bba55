import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand here on CPU; prebuilt on the GPU box)."""
    from latent_diffusion_planning_b200 import _native, build
    if not _native.LIB_PATH.exists():
        build.build()
    return _native.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from latent_diffusion_planning_b200 import _native
    _native.check(_native.load().ldp_device_check())
    return torch.device("cuda:0")
