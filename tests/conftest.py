import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (library, emulator, and -- where the reference tree
    exists -- oracle/_ref)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def cuda(built):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    return torch.device("cuda:0")
