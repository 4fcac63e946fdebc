"""GPU: parity of the CUDA path (through the C ABI of libfftconv_cuda.so) with the
oracle, the reference's golden fixtures and the mathematical definition.
Tolerances are BASELINE.json's: relative L2 <= 1e-5 (float32), <= 1e-12 (float64)."""
import numpy as np
import pytest

from oracle import fftconv_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _t(x, dev):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _check(y, a, k, mode, tag):
    tol = H.TOL[a.dtype]
    y = np.asarray(y)
    assert np.all(np.isfinite(y)), tag
    e_def = O.rel_l2(y, O.conv_exact(a, k, mode))
    e_orc = O.rel_l2(y, O.oaconvolve(a, k, mode))
    assert e_def < tol and e_orc < tol, (tag, e_def, e_orc)


def test_golden_reference_vectors_host_pointers(cuda):
    """Every committed reference output (tests/golden) through the host-pointer API."""
    import fftconv_b200 as fc

    g, names = H.golden_conv_cases()
    for name in names:
        a, k = g[f"{name}__a"], g[f"{name}__k"]
        fn = fc.oaconvolve if name.startswith("oa_") else fc.convolve
        for mode in ("full", "same"):
            got = fn(a, k, mode)
            want = g[f"{name}__{mode}"]
            assert got.shape == want.shape and got.dtype == want.dtype
            assert O.rel_l2(got, want) < H.TOL[a.dtype], (name, mode, O.rel_l2(got, want))


def test_golden_hilbert_vectors(cuda):
    import fftconv_b200 as fc

    g, names = H.golden_hilbert_cases()
    for name in names:
        x = g[f"{name}__x"]
        assert O.rel_l2(fc.hilbert(x), g[f"{name}__env"]) < H.TOL[x.dtype], name


def test_python_binding_cases(cuda):
    """py/test/test_fftconv.py:9-70, same arrays and tolerances."""
    import fftconv_b200 as fc

    a = np.array([1, 2, 3, 4, 5], dtype=np.float64)
    k = np.array([0.2, 0.5, 0.2], dtype=np.float64)
    for fn, fn_ in ((fc.convolve, fc.convolve_), (fc.oaconvolve, fc.oaconvolve_)):
        for mode in ("full", "same"):
            expected = np.convolve(a, k, mode=mode)
            np.testing.assert_allclose(fn(a, k, mode=mode), expected, rtol=1e-5, atol=1e-8)
            out = np.empty(expected.size)
            fn_(a, k, out, mode=mode)
            np.testing.assert_allclose(out, expected, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_reference_gtest_shapes(cuda, dt):
    """test/test_fftconv.cpp:127-218: n = k in 4..94 step 9 (convolve + oaconvolve, Full and
    Same, even K included) and k = 95 with n in {200, 350, 500, 1000, 2000, 3000}; abs
    tolerance 1e-9 / 1e-5 (test_common.hpp:9-18) against direct convolution."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(11)
    atol = 1e-9 if dt == np.float64 else 1e-5
    for size in range(4, 100, 9):
        a, k = rng.standard_normal(size).astype(dt), rng.standard_normal(size).astype(dt)
        for mode in ("full", "same"):
            want = O.conv_exact(a, k, mode)
            for fn in (fc.convolve, fc.oaconvolve):
                got = fn(a, k, mode)
                assert np.max(np.abs(got - want)) < atol * max(1.0, np.max(np.abs(want))), (size, mode, fn.__name__)
    k = rng.standard_normal(95).astype(dt)
    for n in (200, 350, 500, 1000, 2000, 3000):
        a = rng.standard_normal(n).astype(dt)
        for mode in ("full", "same"):
            _check(fc.oaconvolve(a, k, mode), a, k, mode, (n, mode))


def test_test_script_shapes(cuda):
    """test/test_script.cpp:29-34: (1664|2816|2304|4352, 65), Same, both precisions."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(12)
    for dt in (np.float64, np.float32):
        for n in (1664, 2816, 2304, 4352):
            a, k = rng.standard_normal(n).astype(dt), rng.standard_normal(65).astype(dt)
            _check(fc.oaconvolve(a, k, "same"), a, k, "same", (n, dt))
            _check(fc.convolve(a, k, "full"), a, k, "full", (n, dt))


N2048_CASES = [
    # batch, n, K
    (4, 65536, 245), (5, 10000, 245), (1, 100000, 245), (3, 5000, 300), (2, 1804, 245),
    (1, 300, 245), (6, 7000, 410), (2, 4000, 221), (7, 1805, 245), (9, 3607, 256), (64, 65536, 245),
]


@pytest.mark.parametrize("batch,n,klen", N2048_CASES + [(512, 65536, 245), (3, 400000, 245), (130, 9000, 300)])
@pytest.mark.parametrize("mode", ["full", "same"])
def test_n2048_kernel_device_pointers(cuda, batch, n, klen, mode, monkeypatch):
    """The N = 2048 kernel (float32, 221 <= K < 411): ragged last blocks, batches that are not
    multiples of four, fewer rows than lanes (virtual rows), slices that start in mid-row,
    one-block rows."""
    import torch

    import fftconv_b200 as fc

    monkeypatch.setenv("FFTCONV_CUDA_NO_W1K", "1")
    rng = np.random.default_rng(batch * 1000 + n + klen)
    a = rng.standard_normal((batch, n)).astype(np.float32)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(np.float32)
    y = fc.oaconvolve(_t(a, cuda), _t(k, cuda), mode)
    torch.cuda.synchronize()
    _check(y.cpu().numpy(), a, k, mode, (batch, n, klen, mode))


@pytest.mark.parametrize("groups", [1, 2, 3, 4])
def test_n2048_kernel_group_variants(cuda, groups, monkeypatch):
    import torch

    import fftconv_b200 as fc

    monkeypatch.setenv("FFTCONV_CUDA_NO_W1K", "1")
    monkeypatch.setenv("FFTCONV_CUDA_N2048_GROUPS", str(groups))
    rng = np.random.default_rng(groups)
    a = rng.standard_normal((37, 30000)).astype(np.float32)
    k = (rng.standard_normal(245) / np.sqrt(245)).astype(np.float32)
    y = fc.oaconvolve(_t(a, cuda), _t(k, cuda), "full")
    torch.cuda.synchronize()
    _check(y.cpu().numpy(), a, k, "full", groups)


def test_n2048_matches_generic_kernel(cuda, monkeypatch):
    """Two independent CUDA implementations of the same segments agree."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(5)
    a = rng.standard_normal((6, 20000)).astype(np.float32)
    k = (rng.standard_normal(245) / np.sqrt(245)).astype(np.float32)
    y_fast = fc.oaconvolve(_t(a, cuda), _t(k, cuda), "full").cpu().numpy()
    monkeypatch.setenv("FFTCONV_CUDA_FORCE_GENERIC", "1")
    y_gen = fc.oaconvolve(_t(a, cuda), _t(k, cuda), "full").cpu().numpy()
    torch.cuda.synchronize()
    assert O.rel_l2(y_fast, y_gen) < 1e-6
    _check(y_gen, a, k, "full", "generic")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("klen", [3, 6, 7, 20, 35, 64, 65, 119, 165, 220, 411, 768, 769, 1500])
def test_generic_overlap_add_every_table_row(cuda, dt, klen):
    """Every row of the segment-size table (fftconv.hpp:38-49) plus the doubling rule."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(klen)
    n = max(3 * fc.optimal_fft_size(klen) + 17, klen)
    a = rng.standard_normal((3, n)).astype(dt)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(dt)
    for mode in ("full", "same"):
        _check(fc.oaconvolve(a, k, mode), a, k, mode, (klen, dt, mode))


def test_strided_rows_and_in_place_outputs(cuda):
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(8)
    big = rng.standard_normal((5, 9000)).astype(np.float32)
    a = big[:, :8000]                       # row stride 9000, row length 8000
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    ta = _t(big, cuda)[:, :8000]
    out = torch.full((5, 8300), float("nan"), device=cuda)[:, :8244]
    fc.oaconvolve_(ta, _t(k, cuda), out, "full")
    torch.cuda.synchronize()
    _check(out.cpu().numpy(), np.ascontiguousarray(a), k, "full", "strided")
    out_h = np.full((5, 8300), np.nan, dtype=np.float32)
    fc.oaconvolve_(a, k, out_h[:, :8244], "full")
    _check(out_h[:, :8244], np.ascontiguousarray(a), k, "full", "strided-host")
    assert np.all(np.isnan(out_h[:, 8244:]))


def test_linearity_and_shift_properties_at_full_size(cuda):
    """BASELINE config 2 at full size (4096 x 65536, K = 245): dense CPU comparison of a
    row sample plus size-independent properties -- linearity in the signal and the
    response to a unit impulse (= the taps themselves, exactly placed)."""
    import torch

    import fftconv_b200 as fc

    g = torch.Generator(device=cuda).manual_seed(1)
    a = torch.randn((4096, 65536), device=cuda, generator=g)
    k = torch.randn(245, device=cuda, generator=g) / 245 ** 0.5
    y = fc.oaconvolve(a, k, "full")
    rows = [0, 1, 2, 3, 1023, 2048, 4093, 4094, 4095]
    _check(y[rows].cpu().numpy(), a[rows].cpu().numpy(), k.cpu().numpy(), "full", "c2 rows")
    # linearity: conv(2a + b) = 2 conv(a) + conv(b) with b = a rolled by one row
    b = torch.roll(a, 1, 0)
    y2 = fc.oaconvolve(2 * a[:256] + b[:256], k, "full")
    lin = 2 * y[:256] + torch.roll(y, 1, 0)[:256]
    assert float(torch.linalg.norm(y2 - lin) / torch.linalg.norm(lin)) < 1e-5
    # impulse response: a unit impulse at position p reproduces the taps at p..p+K-1
    imp = torch.zeros((4, 65536), device=cuda)
    pos = [0, 1803, 1804, 65535]
    for r, p in enumerate(pos):
        imp[r, p] = 1.0
    yi = fc.oaconvolve(imp, k, "full").cpu().numpy()
    kk = k.cpu().numpy()
    for r, p in enumerate(pos):
        want = np.zeros(65536 + 244, dtype=np.float32)
        want[p:p + 245] = kk
        assert np.max(np.abs(yi[r] - want)) < 2e-6, (r, p)


def test_batched_convolve_config3_sample(cuda):
    """BASELINE config 3 shape (float64, n = 16384, K = 165, Full), 16 rows densely."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(21)
    a = rng.standard_normal((16, 16384))
    k = rng.standard_normal(165) / np.sqrt(165)
    for mode in ("full", "same"):
        got = fc.convolve(a, k, mode)
        assert O.rel_l2(got, O.conv_exact(a, k, mode)) < 1e-12
        assert O.rel_l2(got, O.convolve(a, k, mode)) < 1e-12


def test_hilbert_config4_sample(cuda):
    """BASELINE config 4 shape (float32, 8192-sample A-lines): dense check of 33 lines
    (odd count: exercises the unpaired last row) for noise and an RF-like pulse."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(22)
    t = np.arange(8192)
    x = rng.standard_normal((33, 8192)).astype(np.float32)
    x[1] = (np.exp(-0.5 * ((t - 4000) / 60.0) ** 2) * np.cos(2 * np.pi * 0.11 * t)).astype(np.float32)
    x[2] = np.cos(4 * np.pi * t / 8191).astype(np.float32)   # benchmark/bench_hilbert.cpp:15-17
    env = fc.hilbert(x)
    assert O.rel_l2(env, O.hilbert_exact(x)) < 1e-5
    assert O.rel_l2(env, O.hilbert(x)) < 1e-5
    x64 = x[:5].astype(np.float64)
    assert O.rel_l2(fc.hilbert(x64), O.hilbert_exact(x64)) < 1e-12


@pytest.mark.parametrize("rows,n", [(301, 8192), (2371, 1024), (1190, 2048), (600, 4096), (4737, 512)])
def test_hilbert_packed_persistent_kernel(cuda, rows, n):
    """Batches large enough for the packed (four rows per FFT) persistent TMA-fed kernel, with
    row counts that are not multiples of four; also rows at a 4-byte offset and with a padded
    stride (the one-shot packed kernel), checked densely on a subset of rows."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(rows + n)
    x = rng.standard_normal((rows, n)).astype(np.float32)
    tx = _t(x, cuda)
    env = fc.hilbert(tx)
    torch.cuda.synchronize()
    pick = np.unique(np.concatenate([np.arange(0, rows, max(1, rows // 40)), np.arange(rows - 9, rows)]))
    want = O.hilbert_exact(x[pick])
    assert O.rel_l2(env[pick].cpu().numpy(), want) < 1e-5
    assert torch.isfinite(env).all()
    # misaligned view: rows start 4 bytes off a 16-byte boundary, stride n + 1
    store = torch.zeros(rows * (n + 1) + 1, dtype=torch.float32, device=cuda)
    view = store[1:].view(rows, n + 1)[:, :n]
    view.copy_(tx)
    env2 = torch.empty_like(tx)
    fc.hilbert_(view, env2)
    torch.cuda.synchronize()
    assert O.rel_l2(env2[pick].cpu().numpy(), want) < 1e-5


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 64, 1024, 4096])
def test_hilbert_power_of_two_lengths(cuda, n):
    import fftconv_b200 as fc

    rng = np.random.default_rng(n)
    for dt in (np.float32, np.float64):
        x = rng.standard_normal((3, n)).astype(dt)
        got = fc.hilbert(x)
        want = O.hilbert_exact(x)
        assert np.linalg.norm(got - want) <= H.TOL[np.dtype(dt)] * max(np.linalg.norm(want), 1.0), (n, dt)


@pytest.mark.parametrize("n", [3, 5, 7, 10, 11, 100, 333, 1000, 2047, 2049, 4095])
def test_hilbert_any_length(cuda, n):
    """hilbert.hpp:16-66 accepts any n > 0 (odd included): chirp-z path."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(n)
    for dt in (np.float32, np.float64):
        x = rng.standard_normal((3, n)).astype(dt)
        got = fc.hilbert(x)
        assert O.rel_l2(got, O.hilbert_exact(x)) < H.TOL[np.dtype(dt)], (n, dt, O.rel_l2(got, O.hilbert_exact(x)))
        assert O.rel_l2(got, O.hilbert(x)) < H.TOL[np.dtype(dt)]
    # the reference's known-answer vector (test/test_hilbert.cpp:6-24), abs 1e-6
    for dt in (np.float64, np.float32):
        assert np.max(np.abs(fc.hilbert(H.HILBERT_IN.astype(dt)) - H.HILBERT_EXPECT)) < 1e-6


def test_long_signal_chunked_by_linearity(cuda):
    """Config 5 in miniature: one long row cut into contiguous chunks, each convolved
    in Full mode, K-1 tails added into the neighbour (fftconv_cuda_add_*)."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(31)
    n, klen, parts = 1_000_003, 245, 4
    a = rng.standard_normal(n).astype(np.float32)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(np.float32)
    ta, tk = _t(a, cuda), _t(k, cuda)
    whole = fc.oaconvolve(ta, tk, "full")
    out = torch.zeros(n + klen - 1, device=cuda)
    bounds = [n * i // parts for i in range(parts + 1)]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        y = fc.oaconvolve(ta[lo:hi], tk, "full")
        fc.add_(out[lo:hi + klen - 1], y)
    torch.cuda.synchronize()
    assert float(torch.linalg.norm(out - whole) / torch.linalg.norm(whole)) < 1e-6
    idx = np.concatenate([np.arange(0, 600), np.array(bounds[1:-1])[:, None].repeat(600, 1).ravel()
                          + np.tile(np.arange(-300, 300), parts - 1), np.arange(n + klen - 601, n + klen - 1)])
    want = O.conv_exact_at(a, k, idx)
    got = whole.cpu().numpy()[idx]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5


def test_launch_counter_and_errors(cuda):
    import fftconv_b200 as fc

    before = fc.launch_count()
    fc.oaconvolve(np.ones(100, np.float32), np.ones(5, np.float32))
    assert fc.launch_count() == before + 1   # 5 taps: the direct kernel
    fc.oaconvolve(np.ones(100, np.float32), np.ones(25, np.float32))
    assert fc.launch_count() == before + 3   # spectrum + segment kernel
    with pytest.raises(RuntimeError, match="expected"):
        fc.oaconvolve_(np.ones(100), np.ones(5), np.empty(100), "full")


def test_cpp_headers_dropin(cuda):
    """tests/cpp/test_headers.cpp: the reference's GTest cases restated against
    include/fftconv/*.hpp of this repo (span templates -> C ABI -> CUDA)."""
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "cpp", "test_headers")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr


def test_reference_pybind_module_dropin(cuda):
    """The reference's own py/main.cpp, compiled unchanged against our headers
    (fftconv_b200/refbinding), passes the reference's Python tests
    (py/test/test_fftconv.py:9-70, same arrays and tolerances)."""
    import glob
    import os
    import sys

    from scipy import signal

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "fftconv_b200", "refbinding")
    if not glob.glob(os.path.join(d, "pyfftconv", "_pyfftconv*.so")):
        pytest.skip("reference binding not built (needs /root/reference at build time)")
    sys.path.insert(0, d)
    try:
        import pyfftconv as ref_mod
    finally:
        sys.path.remove(d)
    a = np.array([1, 2, 3, 4, 5], dtype=np.float64)
    k = np.array([0.2, 0.5, 0.2], dtype=np.float64)
    for fn, fn_ in ((ref_mod.convolve, ref_mod.convolve_), (ref_mod.oaconvolve, ref_mod.oaconvolve_)):
        for mode in ("full", "same"):
            expected = np.convolve(a, k, mode=mode)
            np.testing.assert_allclose(fn(a, k, mode=mode), expected, rtol=1e-5, atol=1e-8)
            out = np.empty(expected.size)
            fn_(a, k, out, mode=mode)
            np.testing.assert_allclose(out, expected, rtol=1e-5, atol=1e-8)
    expected = np.abs(signal.hilbert(a))
    np.testing.assert_allclose(ref_mod.hilbert(a), expected, rtol=1e-5, atol=1e-8)
    out = np.empty(a.size)
    ref_mod.hilbert_(a, out)
    np.testing.assert_allclose(out, expected, rtol=1e-5, atol=1e-8)
    # float32 overload and the binding's own error paths (py/main.cpp:21-27, :38-44)
    a32 = np.random.default_rng(0).standard_normal(5000).astype(np.float32)
    k32 = np.random.default_rng(1).standard_normal(245).astype(np.float32)
    got = ref_mod.oaconvolve(a32, k32, mode="same")
    assert got.dtype == np.float32 and O.rel_l2(got, O.conv_exact(a32, k32, "same")) < 1e-5
    with pytest.raises(RuntimeError, match="Unsupported convolution mode"):
        ref_mod.convolve(a, k, mode="valid")
    with pytest.raises(RuntimeError, match="Kernel size must be smaller"):
        ref_mod.convolve(k, a)
    assert ref_mod.__version__ == "0.5.1"


@pytest.mark.parametrize("dt,n,klen", [(np.float64, 16384, 165), (np.float32, 100_000, 1000), (np.float64, 60_000, 33),
                                       (np.float32, 200_000, 245), (np.float32, 16400, 16000)])
def test_convolve_one_long_fft(cuda, dt, n, klen, monkeypatch):
    """convolve_fftw semantics for rows longer than one CTA's FFT: ONE transform of length
    nextpow2(n + K - 1) = 16 x N_sub through global scratch (long_fft_kernels.cuh); checked
    against the definition, the oracle, and the segment-FFT route (same linear convolution)."""
    import fftconv_b200 as fc

    rng = np.random.default_rng(n + klen)
    a = rng.standard_normal((3, n)).astype(dt)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(dt)
    tol = H.TOL[np.dtype(dt)]
    monkeypatch.setenv("FFTCONV_CUDA_CONV_SINGLE_FFT", "1")
    before = fc.launch_count()
    for mode in ("full", "same"):
        got = fc.convolve(a, k, mode)
        assert O.rel_l2(got, O.conv_exact(a, k, mode)) < tol, (dt, n, klen, mode)
        assert O.rel_l2(got, O.convolve(a, k, mode)) < tol
    launches_long = fc.launch_count() - before
    assert launches_long >= 2 * 4   # spectrum + outer fwd + inner + outer inv, per call
    # the default beyond one CTA's FFT: segment FFTs where the taps allow (two launches per call)
    monkeypatch.delenv("FFTCONV_CUDA_CONV_SINGLE_FFT")
    before = fc.launch_count()
    alt = fc.convolve(a, k, "full")
    assert O.rel_l2(alt, O.conv_exact(a, k, "full")) < tol
    assert fc.launch_count() - before == (2 if klen <= 4096 else 4)


@pytest.mark.parametrize("dt,n", [(np.float32, 32768), (np.float32, 262144), (np.float64, 16384), (np.float64, 131072)])
def test_hilbert_long_lines(cuda, dt, n):
    import fftconv_b200 as fc

    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(dt)
    got = fc.hilbert(x)
    assert O.rel_l2(got, O.hilbert_exact(x)) < H.TOL[np.dtype(dt)]


def test_cuda_array_interface_inputs(cuda):
    """device arrays of other libraries are taken in place through
    __cuda_array_interface__ (exercised with a minimal wrapper around torch storage)."""
    import torch

    import fftconv_b200 as fc

    class CAI:
        def __init__(self, t):
            self.t = t
            self.__cuda_array_interface__ = {
                "shape": tuple(t.shape), "typestr": "<f4" if t.dtype == torch.float32 else "<f8",
                "data": (t.data_ptr(), False), "version": 3, "strides": None, "stream": 1}

    rng = np.random.default_rng(77)
    a = rng.standard_normal((6, 9000)).astype(np.float32)
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    ta, tk = _t(a, cuda), _t(k, cuda)
    out = torch.empty((6, 9244), device=cuda)
    torch.cuda.synchronize()
    fc.oaconvolve_(CAI(ta), CAI(tk), CAI(out), "full")
    torch.cuda.synchronize()
    _check(out.cpu().numpy(), a, k, "full", "cai")
    env = torch.empty((6, 8192), device=cuda)
    fc.hilbert_(CAI(ta[:, :8192].contiguous()), CAI(env))
    torch.cuda.synchronize()
    assert O.rel_l2(env.cpu().numpy(), O.hilbert_exact(a[:, :8192])) < 1e-5


def test_dlpack_inputs(cuda):
    """Any DLPack exporter on a CUDA device is taken in place (a minimal wrapper that exposes only
    __dlpack__ / __dlpack_device__ stands in for JAX / CuPy arrays)."""
    import torch

    import fftconv_b200 as fc

    class DL:
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, *a, **kw):
            return self.t.__dlpack__(*a, **kw)

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    rng = np.random.default_rng(78)
    a = rng.standard_normal((5, 7000))
    k = rng.standard_normal(165) / 13
    ta, tk = _t(a, cuda), _t(k, cuda)
    out = torch.full((5, 7164), float("nan"), dtype=torch.float64, device=cuda)
    fc.oaconvolve_(DL(ta), DL(tk), DL(out), "full")
    torch.cuda.synchronize()
    _check(out.cpu().numpy(), a, k, "full", "dlpack")


@pytest.mark.parametrize("dt,klen", [(np.float32, 245), (np.float32, 65), (np.float64, 165), (np.float32, 7)])
def test_filter_handle_reuses_the_spectrum(cuda, dt, klen):
    """a filter handle gives the same samples as the plain call (to rounding: a handle
    always uses the large-job block size, a small plain call the reference's table) and launches no
    spectrum kernel."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(klen)
    a = rng.standard_normal((5, 20000)).astype(dt)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(dt)
    ta = _t(a, cuda)
    f = fc.Filter(k)
    for mode in ("full", "same"):
        want = fc.oaconvolve(ta, _t(k, cuda), mode)
        torch.cuda.synchronize()
        before = fc.launch_count()
        got = f.oaconvolve(ta, mode)
        torch.cuda.synchronize()
        assert fc.launch_count() - before == 1            # the overlap-add kernel alone
        assert O.rel_l2(got.cpu().numpy(), want.cpu().numpy()) < H.TOL[np.dtype(dt)]
        _check(got.cpu().numpy(), a, k, mode, ("filter", klen, mode))
    # host arrays go through the same handle
    _check(f.oaconvolve(a, "full"), a, k, "full", ("filter-host", klen))
    f.close()


@pytest.mark.parametrize("dt,klen,pieces", [
    (np.float32, 245, [5000, 1804, 100, 1, 30000, 244, 245, 7]),   # pieces shorter than K-1 too
    (np.float64, 165, [4096, 4096, 333]),
    (np.float32, 20, [64, 1, 1, 1000]),
])
def test_streaming_overlap_add_across_calls(cuda, dt, klen, pieces):
    """Pieces pushed one at a time + flush == oaconvolve of the whole signal (Full)."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(sum(pieces))
    rows, n = 3, sum(pieces)
    a = rng.standard_normal((rows, n)).astype(dt)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(dt)
    ta = _t(a, cuda)
    s = fc.Filter(k).stream(rows)
    outs, pos = [], 0
    for p in pieces:
        outs.append(s.push(ta[:, pos:pos + p]))     # row-strided views: no copy
        pos += p
    outs.append(s.flush(ta))
    got = torch.cat(outs, dim=1).cpu().numpy()
    assert got.shape == (rows, n + klen - 1)
    _check(got, a, k, "full", ("stream", klen))
    # the state is reset by flush: a second pass gives the same answer
    again = torch.cat([s.push(ta), s.flush(ta)], dim=1).cpu().numpy()
    assert O.rel_l2(again, got) < H.TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("klen", [1, 3, 7, 8, 9, 16, 17, 20, 65, 165, 220, 221, 300, 360, 361, 410, 411, 500, 560, 561, 640,
                                  641, 768, 1024, 1025])
def test_large_jobs_use_the_throughput_segment_size(cuda, dt, klen, monkeypatch):
    """Jobs of >= 4 Mi samples move short kernels off the reference's (CPU cost model) segment
    size onto the fastest kernel's; the samples are the same linear convolution either way."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(klen)
    a = rng.standard_normal((64 if klen < 400 else 67, 65536)).astype(dt)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(dt)
    ta, tk = _t(a, cuda), _t(k, cuda)
    tol = H.TOL[np.dtype(dt)]
    for mode in ("full", "same"):
        before = fc.launch_count()
        y = fc.oaconvolve(ta, tk, mode)
        if klen <= (16 if dt == np.float32 else 8):
            assert fc.launch_count() - before == 1     # direct time-domain kernel
        elif dt == np.float32 and klen <= 360:
            assert fc.launch_count() - before == 1     # the warp-per-FFT kernel transforms the taps itself
        elif klen <= (1024 if dt == np.float32 else 640):
            assert fc.launch_count() - before == 2     # spectrum + the N = 2048 kernel (wide form beyond 410 taps)
        monkeypatch.setenv("FFTCONV_CUDA_STRICT_TABLE", "1")
        y_table = fc.oaconvolve(ta, tk, mode)
        monkeypatch.delenv("FFTCONV_CUDA_STRICT_TABLE")
        torch.cuda.synchronize()
        assert O.rel_l2(y.cpu().numpy(), y_table.cpu().numpy()) < tol
        pick = [0, 1, 31, 63]
        assert O.rel_l2(y[pick].cpu().numpy(), O.conv_exact(a[pick], k, mode)) < tol, (klen, mode)


# ---- envelope with fused post-processing --------------------------
ENV_CASES = [
    # (dtype, rows, n, device tensors, what it exercises)
    (np.float32, 1300, 8192, True),    # packed persistent TMA-fed kernel (config 4 shape)
    (np.float32, 37, 8192, True),      # two rows per FFT (small batch)
    (np.float32, 5, 64, False),        # small power of two, host pointers
    (np.float32, 6, 1000, False),      # chirp-z
    (np.float32, 3, 32768, True),      # one long FFT through global scratch
    (np.float64, 9, 4096, True),
    (np.float64, 4, 777, False),
]


@pytest.mark.parametrize("dt,rows,n,on_device", ENV_CASES)
@pytest.mark.parametrize("dec,log", [(1, True), (3, False), (4, True), (7, True)])
def test_envelope_epilogue_matches_oracle(cuda, dt, rows, n, on_device, dec, log):
    """fftconv_cuda_envelope_*: decimation and log compression in the store of every Hilbert
    kernel, against the oracle restatement (reference hilbert + the stated post-processing)."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(n + 13 * dec + rows)
    x = (rng.standard_normal((rows, n)) * np.exp(rng.uniform(-4, 1, (rows, 1)))).astype(dt)
    kw = dict(log_compress=log, ref=1.5, dynamic_range_db=50.0, decimate=dec)
    if on_device:
        got = fc.envelope(_t(x, cuda), **kw)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
    else:
        got = fc.envelope(x, **kw)
    pick = np.unique(np.concatenate([np.arange(0, rows, max(1, rows // 24)), np.arange(max(0, rows - 5), rows)]))
    want = O.envelope(x[pick], **kw)
    assert got.shape == (rows, (n + dec - 1) // dec) and got.dtype == x.dtype
    assert np.all(np.isfinite(got))
    if log:
        assert got.min() >= 0.0 and got.max() <= 1.0
        # absolute: outputs live in [0, 1]; 1e-5 / 1e-12 of full scale per sample on average
        err = np.linalg.norm(got[pick].astype(np.float64) - want) / np.sqrt(want.size)
        assert err < H.TOL[np.dtype(dt)] * 4, err
    else:
        assert O.rel_l2(got[pick], want) < H.TOL[np.dtype(dt)]


def test_envelope_plain_options_equal_hilbert(cuda):
    import fftconv_b200 as fc

    rng = np.random.default_rng(5)
    for dt, n in ((np.float32, 2048), (np.float64, 300)):
        x = rng.standard_normal((7, n)).astype(dt)
        assert np.array_equal(fc.envelope(x), fc.hilbert(x))
    with pytest.raises(RuntimeError):
        fc.envelope_(x, np.empty((7, 300), dtype=np.float64), decimate=2)   # wrong output length
    with pytest.raises(RuntimeError):
        fc.envelope(x, log_compress=True, ref=0.0)


@pytest.mark.parametrize("dt,rows,n,klen,on_device", [(np.float32, 600, 8192, 245, True), (np.float32, 9, 5000, 65, False),
                                                      (np.float64, 12, 4096, 95, True)])
def test_rf_envelope_chain(cuda, dt, rows, n, klen, on_device):
    """Band-pass (oaconvolve Same) then envelope without leaving the device, against the oracle
    chain and against the two separate public calls."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(klen + rows)
    x = rng.standard_normal((rows, n)).astype(dt)
    t = np.arange(klen) - klen // 2
    k = (np.hamming(klen) * np.cos(2 * np.pi * 0.12 * t)).astype(dt)
    k /= np.abs(k).sum()
    kw = dict(log_compress=True, ref=0.2, dynamic_range_db=45.0, decimate=2)
    if on_device:
        tx, tk = _t(x, cuda), _t(k, cuda)
        got = fc.rf_envelope(tx, tk, **kw)
        lin = fc.rf_envelope(tx, tk)
        two = fc.hilbert(fc.oaconvolve(tx, tk, "same"))
        torch.cuda.synchronize()
        got, lin, two = got.cpu().numpy(), lin.cpu().numpy(), two.cpu().numpy()
    else:
        got, lin = fc.rf_envelope(x, k, **kw), fc.rf_envelope(x, k)
        two = fc.hilbert(fc.oaconvolve(x, k, "same"))
    assert np.array_equal(lin, two)
    pick = np.arange(0, rows, max(1, rows // 16))
    tol = H.TOL[np.dtype(dt)]
    assert O.rel_l2(lin[pick], O.rf_envelope(x[pick], k)) < tol * 3   # two chained transforms
    want = O.rf_envelope(x[pick], k, **kw)
    assert np.linalg.norm(got[pick].astype(np.float64) - want) / np.sqrt(want.size) < tol * 4


@pytest.mark.parametrize("dt,n,klen", [(np.float32, 30000, 10000), (np.float32, 9000, 8999), (np.float64, 20000, 5000),
                                       (np.float64, 70000, 4097)])
def test_kernels_longer_than_a_segment_fft(cuda, dt, n, klen):
    """The reference keeps doubling the segment FFT for long kernels (fftconv.hpp:61-65); here the
    taps are cut into blocks and the partial convolutions summed in order: same linear
    convolution, Full and Same, host and device pointers, oaconvolve and convolve."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(klen)
    a = rng.standard_normal((3, n)).astype(dt)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(dt)
    tol = H.TOL[np.dtype(dt)]
    for mode in ("full", "same"):
        want = O.conv_exact(a, k, mode)
        got = fc.oaconvolve(a, k, mode)
        assert O.rel_l2(got, want) < tol, (mode, O.rel_l2(got, want))
        tgot = fc.oaconvolve(_t(a, cuda), _t(k, cuda), mode)
        torch.cuda.synchronize()
        assert np.array_equal(tgot.cpu().numpy(), got)
    assert O.rel_l2(fc.convolve(a[0], k, "full"), O.conv_exact(a[0], k, "full")) < tol


N2048_F64_CASES = [
    # batch, n, K
    (4, 65536, 245), (5, 10000, 245), (1, 100000, 245), (3, 5000, 300), (2, 1804, 245), (1, 300, 245),
    (6, 7000, 410), (2, 4000, 221), (7, 1805, 245), (9, 3607, 256), (33, 40000, 245), (1024, 16384, 165),
]


@pytest.mark.parametrize("batch,n,klen", N2048_F64_CASES)
@pytest.mark.parametrize("mode", ["full", "same"])
def test_n2048_kernel_float64(cuda, batch, n, klen, mode, monkeypatch):
    """oa_n2048_kernel<double, ...>: the same engine with one float64 transform per thread (two
    streams per group): ragged last segments, odd batches, rows cut into chunks, rows at an
    8-byte offset; config 3's shape (1024 x 16384, K = 165) rides on it through the throughput
    rule.  Dense check on a subset of rows for the large case; bit-equality with the register-
    blocked engine is not expected, 1e-12 relative L2 with the definition is."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(batch * 1000 + n + klen)
    a = rng.standard_normal((batch, n))
    k = rng.standard_normal(klen) / np.sqrt(klen)
    ta, tk = _t(a, cuda), _t(k, cuda)
    before = fc.launch_count()
    y = fc.oaconvolve(ta, tk, mode) if batch < 1000 else fc.convolve(ta, tk, mode)
    torch.cuda.synchronize()
    assert fc.launch_count() - before == 2          # spectrum + one fused kernel
    pick = np.arange(batch) if batch < 64 else np.unique(np.r_[np.arange(0, batch, 97), batch - 3, batch - 2, batch - 1])
    _check(y[pick].cpu().numpy(), a[pick], k, mode, (batch, n, klen, mode))
    assert torch.isfinite(y).all()
    # the register-blocked engine on the same input
    monkeypatch.setenv("FFTCONV_CUDA_NO_N2048_F64", "1")
    y2 = fc.oaconvolve(ta, tk, mode)
    torch.cuda.synchronize()
    assert O.rel_l2(y[pick].cpu().numpy(), y2[pick].cpu().numpy()) < 1e-13
    monkeypatch.delenv("FFTCONV_CUDA_NO_N2048_F64")
    # misaligned rows: 8 bytes off a 16-byte boundary, odd stride
    if batch <= 9:
        store = torch.zeros(batch * (n + 1) + 1, dtype=torch.float64, device=cuda)
        view = store[1:].view(batch, n + 1)[:, :n]
        view.copy_(ta)
        out_len = y.shape[-1]
        ostore = torch.full((batch * (out_len + 3) + 1,), float("nan"), dtype=torch.float64, device=cuda)
        oview = ostore[1:].view(batch, out_len + 3)[:, :out_len]
        fc.oaconvolve_(view, tk, oview, mode)
        torch.cuda.synchronize()
        assert torch.equal(oview, y)
        assert torch.isnan(ostore[1:].view(batch, out_len + 3)[:, out_len:]).all()   # nothing written past a row


@pytest.mark.parametrize("dt,n", [(np.float32, 10000), (np.float32, 12345), (np.float32, 50001), (np.float32, 131072 - 3),
                                  (np.float64, 5000), (np.float64, 6001), (np.float64, 40000)])
def test_hilbert_long_lines_of_any_length(cuda, dt, n):
    """hilbert.hpp:16-66 accepts any n: lines that are not a power of two and whose chirp-z FFT
    (M >= 2 n - 1) exceeds a CTA's shared memory run the chirp filter's two convolutions as split
    FFTs through global scratch (long_fft_kernels.cuh).  Odd row count: unpaired last row."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(dt)
    t = np.arange(n)
    x[1] = (np.exp(-0.5 * ((t - n // 3) / 200.0) ** 2) * np.cos(2 * np.pi * 0.07 * t)).astype(dt)
    want = O.hilbert_exact(x)
    got = fc.hilbert(x)
    tol = H.TOL[np.dtype(dt)]
    assert O.rel_l2(got, want) < tol, (n, dt, O.rel_l2(got, want))
    tx = _t(x, cuda)
    tgot = fc.hilbert(tx)
    dec = fc.envelope(tx, decimate=5, log_compress=True, ref=1.0, dynamic_range_db=50.0)
    torch.cuda.synchronize()
    assert np.array_equal(tgot.cpu().numpy(), got)
    wdec = O.envelope(x, log_compress=True, ref=1.0, dynamic_range_db=50.0, decimate=5)
    assert np.linalg.norm(dec.cpu().numpy().astype(np.float64) - wdec) / np.sqrt(wdec.size) < tol * 4


DIRECT_CASES = [
    # batch, n, K
    (1, 1, 1), (1, 5, 3), (3, 100, 5), (2, 4352, 8), (5, 1023, 7), (4, 1024, 8), (3, 1025, 2), (7, 5000, 4),
    (2, 70001, 6), (64, 65536, 3), (33, 40000, 8), (1, 3000000, 5), (3, 4352, 16), (5, 9000, 9), (6, 4100, 13),
]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("batch,n,klen", DIRECT_CASES)
def test_direct_kernel_for_short_kernels(cuda, dt, batch, n, klen, monkeypatch):
    """Kernels of at most 16 (float32) / 8 (float64) taps are evaluated in the time domain (direct_kernels.cuh): against the
    definition, against the FFT path on the same input (FFTCONV_CUDA_NO_DIRECT=1), Full and Same,
    host and device pointers, rows at every 4-byte / 8-byte alignment (odd output strides), and
    nothing written past a row."""
    import torch

    import fftconv_b200 as fc

    if klen > n:
        pytest.skip("binding requires n >= K")
    rng = np.random.default_rng(batch * 131 + n + klen)
    a = rng.standard_normal((batch, n)).astype(dt)
    k = rng.standard_normal(klen).astype(dt)
    tol = H.TOL[np.dtype(dt)]
    ta, tk = _t(a, cuda), _t(k, cuda)
    for mode in ("full", "same"):
        want = O.conv_exact(a, k, mode)
        y = fc.oaconvolve(ta, tk, mode)
        torch.cuda.synchronize()
        assert O.rel_l2(y.cpu().numpy(), want) < tol, (mode, O.rel_l2(y.cpu().numpy(), want))
        if n <= 70001:
            assert O.rel_l2(fc.convolve(a, k, mode), want) < tol       # host pointers, convolve entry
        monkeypatch.setenv("FFTCONV_CUDA_NO_DIRECT", "1")
        y_fft = fc.oaconvolve(ta, tk, mode)
        monkeypatch.delenv("FFTCONV_CUDA_NO_DIRECT")
        torch.cuda.synchronize()
        assert O.rel_l2(y.cpu().numpy(), y_fft.cpu().numpy()) < tol * 2
        # misaligned rows in and out, NaN guard band after every output row
        if batch <= 7:
            out_len = y.shape[-1]
            for off in (1, 2, 3):
                store = torch.zeros(batch * (n + 3) + off, dtype=ta.dtype, device=cuda)
                view = store[off:].view(batch, n + 3)[:, :n]
                view.copy_(ta)
                ostore = torch.full((batch * (out_len + 5) + off,), float("nan"), dtype=ta.dtype, device=cuda)
                full_view = ostore[off:].view(batch, out_len + 5)
                fc.oaconvolve_(view, tk, full_view[:, :out_len], mode)
                torch.cuda.synchronize()
                assert torch.equal(full_view[:, :out_len], y), (mode, off)
                assert torch.isnan(full_view[:, out_len:]).all()
                assert torch.isnan(ostore[:off]).all()


def test_small_jobs_take_one_launch_and_match_the_reference_cases(cuda, monkeypatch):
    """Default dispatch for small jobs (<= 4 Mi multiply-adds, <= 256 taps): direct_small_kernel, one
    launch.  Replays the reference's own small cases through it: golden vectors (tests/golden), the
    Python binding cases (py/test/test_fftconv.py:9-70), the GTest shapes (test/test_fftconv.cpp:93-218)
    and the test_script sizes (test/test_script.cpp:29-34, config 1)."""
    import torch

    import fftconv_b200 as fc

    monkeypatch.delenv("FFTCONV_CUDA_NO_SMALL_DIRECT", raising=False)
    g, names = H.golden_conv_cases()
    for name in names:
        a, k = g[f"{name}__a"], g[f"{name}__k"]
        fn = fc.oaconvolve if name.startswith("oa_") else fc.convolve
        for mode in ("full", "same"):
            before = fc.launch_count()
            got = fn(a, k, mode)
            if a.shape[-1] * k.size * (1 if a.ndim == 1 else a.shape[0]) <= 2 ** 21 and k.size <= 256:
                assert fc.launch_count() - before == 1, name
            assert O.rel_l2(got, g[f"{name}__{mode}"]) < H.TOL[a.dtype], (name, mode)
    a = np.array([1, 2, 3, 4, 5], dtype=np.float64)
    k = np.array([0.2, 0.5, 0.2], dtype=np.float64)
    for mode in ("full", "same"):
        assert np.allclose(fc.convolve(a, k, mode), np.convolve(a, k, mode=mode), rtol=1e-5, atol=1e-8)
        assert np.allclose(fc.oaconvolve(a, k, mode), np.convolve(a, k, mode=mode), rtol=1e-5, atol=1e-8)
    rng = np.random.default_rng(9)
    for dt in (np.float64, np.float32):
        for n, klen in ((1000, 95), (2000, 95), (3000, 95), (200, 95), (350, 95), (500, 95), (4, 4), (13, 13), (94, 94),
                        (1664, 65), (2816, 65), (2304, 65), (4352, 65), (4352, 256), (300, 245)):
            a, k = rng.standard_normal(n).astype(dt), rng.standard_normal(klen).astype(dt)
            for mode in ("full", "same"):
                before = fc.launch_count()
                y = fc.oaconvolve(_t(a, cuda), _t(k, cuda), mode)
                torch.cuda.synchronize()
                assert fc.launch_count() - before == 1
                _check(y.cpu().numpy(), a, k, mode, (n, klen, dt, mode))
                _check(fc.convolve(a, k, mode), a, k, mode, ("host", n, klen, dt, mode))
    # a batch with padded strides and a guard band
    a = rng.standard_normal((7, 900)).astype(np.float32)
    k = rng.standard_normal(33).astype(np.float32)
    store = torch.full((7, 940), float("nan"), device=cuda)
    fc.oaconvolve_(_t(a, cuda), _t(k, cuda), store[:, :932], "full")
    torch.cuda.synchronize()
    _check(store[:, :932].cpu().numpy(), a, k, "full", "strided")
    assert torch.isnan(store[:, 932:]).all()


def test_small_jobs_rule_boundaries(cuda, monkeypatch):
    """Just above the limits (work, taps) the FFT path serves the call; both agree."""
    import fftconv_b200 as fc

    monkeypatch.delenv("FFTCONV_CUDA_NO_SMALL_DIRECT", raising=False)
    rng = np.random.default_rng(10)
    for n, klen, launches in ((16000, 257, 2), (70000, 65, 2), (4352, 65, 1)):
        a, k = rng.standard_normal(n).astype(np.float32), rng.standard_normal(klen).astype(np.float32)
        before = fc.launch_count()
        y = fc.oaconvolve(a, k, "full")
        assert fc.launch_count() - before == launches, (n, klen)
        _check(y, a, k, "full", (n, klen))


def test_concurrent_callers(cuda):
    """Any thread may call any entry point (include/fftconv_cuda.h): four Python threads (ctypes
    releases the GIL) mixing host-pointer and device-pointer calls, float32 and float64, overlap-add,
    Hilbert and the envelope options; every result is checked."""
    import threading

    import torch

    import fftconv_b200 as fc

    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for it in range(6):
                dt = np.float32 if (seed + it) % 2 == 0 else np.float64
                tol = H.TOL[np.dtype(dt)]
                a = rng.standard_normal((3 + seed, 3000 + 500 * it)).astype(dt)
                k = (rng.standard_normal(20 + 60 * seed) / 8).astype(dt)
                if it % 2 == 0:
                    y = fc.oaconvolve(a, k, "full")
                else:
                    y = fc.oaconvolve(_t(a, cuda), _t(k, cuda), "full").cpu().numpy()
                assert O.rel_l2(y, O.conv_exact(a, k, "full")) < tol
                x = rng.standard_normal((5, 1024)).astype(dt)
                assert O.rel_l2(fc.hilbert(x), O.hilbert_exact(x)) < tol
                e = fc.envelope(x, decimate=2 + seed)
                assert O.rel_l2(e, O.envelope(x, decimate=2 + seed)) < tol
        except Exception as exc:  # noqa: BLE001 - reported below with the thread's seed
            errors.append((seed, repr(exc)))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    assert not errors, errors
