import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without CUDA must fail loudly, not skip silently.
    pass


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    assert torch.cuda.is_available(), "gpu-marked test needs a CUDA device"
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)
