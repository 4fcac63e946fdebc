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


@pytest.fixture(autouse=True)
def _fft_path_for_small_jobs(monkeypatch):
    """The library evaluates SMALL jobs (<= 4 Mi multiply-adds) directly, one thread per output
    (direct_small_kernel).  Most parity tests use small shapes precisely to exercise the FFT kernels
    cheaply, so they switch that rule off; the tests of the rule itself
    (tests/test_gpu_parity.py::test_small_jobs_*) switch it back on."""
    monkeypatch.setenv("FFTCONV_CUDA_NO_SMALL_DIRECT", "1")
