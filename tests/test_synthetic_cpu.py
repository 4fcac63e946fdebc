"""CPU: the counter-based synthetic signal (measurement plumbing).  The numpy twin
is pinned to a HOST build of the very header the device kernel is compiled from; the exact checker
built on it is pinned to np.convolve; the C ABI's chunk / row bounds are pinned to the Python ones."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_gen(built):
    exe = os.path.join(ROOT, "tests", "cpu", "synthetic_host")
    src = exe + ".cpp"
    hdr = os.path.join(ROOT, "fftconv_b200", "csrc", "synthetic.cuh")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    return exe


@pytest.mark.parametrize("dt,name", [(np.float32, "f32"), (np.float64, "f64")])
def test_numpy_twin_matches_the_header(host_gen, dt, name):
    from fftconv_b200 import synthetic as syn

    for seed, first, gain in ((syn.SEED_SIGNAL, 0, 1.0), (syn.SEED_TAPS, 2 ** 40 + 17, 1.0 / np.sqrt(245)),
                              (12345, 2 ** 31 - 100, 0.3)):
        out = subprocess.run([host_gen, str(seed), str(first), "300", repr(float(gain)), name],
                             capture_output=True, text=True, check=True).stdout.split()
        want = np.array([float(v) for v in out]).astype(dt)
        got = syn.signal(seed, first, 300, dt, gain)
        assert np.array_equal(got, want), (seed, first)


def test_signal_statistics():
    from fftconv_b200 import synthetic as syn

    x = syn.signal(syn.SEED_SIGNAL, 0, 1 << 20, np.float64)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3 and np.abs(x).max() < 3.47
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 5e-3          # consecutive counters are uncorrelated


def test_conv_at_is_the_definition():
    from fftconv_b200 import synthetic as syn

    n, K = 5000, 245
    taps = syn.signal(syn.SEED_TAPS, 0, K, np.float32, gain=1 / np.sqrt(K))
    x = syn.signal(syn.SEED_SIGNAL, 1000, n, np.float32)
    full = np.convolve(x.astype(np.float64), taps.astype(np.float64))
    idx = np.array([0, 1, 243, 244, 245, 2500, n - 1, n, n + K - 2])
    got = syn.conv_at(syn.SEED_SIGNAL, n, taps, idx, row_first=1000)
    assert np.allclose(got, full[idx], rtol=0, atol=1e-12)
    idx = syn.cut_window_indices([1000, 3000], n + K - 1, K, extra_random=50, tail=100)
    assert idx.min() == 0 and idx.max() == n + K - 2 and np.all(np.diff(idx) > 0)
    for c in (1000, 3000):
        assert np.all(np.isin(np.arange(c - K, c + K), idx))
    assert np.allclose(syn.conv_at(syn.SEED_SIGNAL, n, taps, idx, row_first=1000), full[idx], rtol=0, atol=1e-12)


def test_c_abi_bounds_match_the_python_partition(built):
    from fftconv_b200 import _capi
    from fftconv_b200 import multigpu as mg

    L = _capi.lib()
    lo, hi = C.c_size_t(), C.c_size_t()
    for n in (2 ** 31, 2 ** 31 + 4103, 1000, 12345678901):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                L.fftconv_cuda_mgpu_chunk_bounds(n, world, r, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == mg.chunk_bounds(n, world, r)
    for batch in (4096, 37, 5, 1):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                L.fftconv_cuda_mgpu_row_bounds(batch, world, r, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == mg.shard_rows(batch, world, r)


def test_mgpu_without_cuda_fails_loudly(built):
    import torch

    from fftconv_b200 import _capi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert _capi.lib().fftconv_cuda_mgpu_init(1, None, C.byref(h)) != 0
    assert b"cuda" in _capi.lib().fftconv_cuda_last_error().lower()
