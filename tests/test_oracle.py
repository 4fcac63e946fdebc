"""CPU: pins the oracle (oracle/fftconv_oracle.py) and the reference-over-shim build
(oracle/_ref) against the reference's literal known-answer vectors, against each
other through the committed golden fixtures, and against the mathematical
definition (the reference tests' own oracle, test/test_fftconv.cpp:93-218)."""
import numpy as np
import pytest

from oracle import fftconv_oracle as O
from oracle import ref
from tests import helpers as H

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")


def test_segment_size_table_matches_reference_rows():
    # include/fftconv/fftconv.hpp:38-66 / optimal_fft_size.csv
    expect = {1: 16, 6: 16, 7: 32, 11: 32, 12: 64, 20: 64, 21: 128, 35: 128, 36: 256, 64: 256,
              65: 512, 95: 512, 119: 512, 120: 1024, 165: 1024, 220: 1024, 221: 2048, 245: 2048,
              410: 2048, 411: 4096, 768: 4096, 769: 8192, 4096: 8192, 4097: 16384, 10000: 32768}
    for k, n in expect.items():
        assert O.get_optimal_fft_size(k) == n, k


def test_cost_model_rederives_the_csv():
    # cost_func.py:5-34 with the saturating-cast rule
    assert tuple(O.cost_model_table()) == O.OPTIMAL_OA_FFT_SIZE


def test_hilbert_known_answer_vector():
    # test/test_hilbert.cpp:6-24, tolerance 1e-6 for double and float
    for dt in (np.float64, np.float32):
        got = O.hilbert(H.HILBERT_IN.astype(dt))
        assert np.max(np.abs(got - H.HILBERT_EXPECT)) < 1e-6


@needs_ref
def test_reference_build_hilbert_known_answer_vector():
    for dt in (np.float64, np.float32):
        got = ref.hilbert(H.HILBERT_IN.astype(dt))
        assert np.max(np.abs(got - H.HILBERT_EXPECT)) < 1e-6


@needs_ref
def test_shim_fft_matches_reference_dft20_vectors():
    # test/test_fftw.cpp:347-393: r2c known answers (abs 1e-7) and c2r(r2c(x))/n == x
    spec = ref.rfft(H.DFT20_IN)
    assert np.max(np.abs(spec.real - H.DFT20_RE)) < 1e-7
    assert np.max(np.abs(spec.imag - H.DFT20_IM)) < 1e-7
    back = ref.irfft_unnormalised(spec, 20) / 20
    assert np.max(np.abs(back - H.DFT20_IN)) < 1e-7
    # and the oracle's FFT (pocketfft) obeys the same sign / scaling convention
    spec_np = np.fft.rfft(H.DFT20_IN)
    assert np.max(np.abs(spec_np.real - H.DFT20_RE)) < 1e-7
    assert np.max(np.abs(spec_np.imag - H.DFT20_IM)) < 1e-7


@needs_ref
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 20, 64, 65, 127, 197, 1000, 2048, 16548])
def test_shim_fft_against_numpy(n):
    rng = np.random.default_rng(n)
    for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        x = rng.standard_normal(n).astype(dt)
        assert O.rel_l2(np.abs(ref.rfft(x) - np.fft.rfft(x.astype(np.float64))), 0 * x[: n // 2 + 1]) >= 0
        err = np.linalg.norm(ref.rfft(x) - np.fft.rfft(x.astype(np.float64))) / np.linalg.norm(x) / np.sqrt(n)
        assert err < tol, (n, dt, err)


def test_oracle_against_golden_reference_outputs():
    g, names = H.golden_conv_cases()
    for name in names:
        a, k = g[f"{name}__a"], g[f"{name}__k"]
        fn = O.oaconvolve if name.startswith("oa_") else O.convolve
        for mode in ("full", "same"):
            want = g[f"{name}__{mode}"]
            got = fn(a, k, mode)
            assert got.shape == want.shape and got.dtype == want.dtype
            assert O.rel_l2(got, want) < H.TOL[a.dtype], (name, mode)
            # and both agree with the definition
            assert O.rel_l2(want, O.conv_exact(a, k, mode)) < H.TOL[a.dtype], (name, mode)
    g, names = H.golden_hilbert_cases()
    for name in names:
        x = g[f"{name}__x"]
        assert O.rel_l2(O.hilbert(x), g[f"{name}__env"]) < H.TOL[x.dtype], name
        assert O.rel_l2(g[f"{name}__env"], O.hilbert_exact(x)) < H.TOL[x.dtype], name


@needs_ref
def test_golden_fixtures_are_current():
    """The committed fixtures are what the reference build produces today."""
    g, names = H.golden_conv_cases()
    for name in names:
        a, k = g[f"{name}__a"], g[f"{name}__k"]
        fn = ref.oaconvolve if name.startswith("oa_") else ref.convolve
        for mode in ("full", "same"):
            np.testing.assert_array_equal(fn(a, k, mode), g[f"{name}__{mode}"])


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_reference_test_shapes_against_definition(dt):
    """Restates test/test_fftconv.cpp:93-218 (arma::conv oracle, n = k in 4..94 step 9;
    k = 95 with n in {200,350,500} and {1000,2000,3000}) on the oracle."""
    rng = np.random.default_rng(3)
    tol = H.TOL[np.dtype(dt)]
    for size in range(4, 100, 9):
        a, k = rng.standard_normal(size).astype(dt), rng.standard_normal(size).astype(dt)
        for mode in ("full", "same"):
            assert O.rel_l2(O.convolve(a, k, mode), O.conv_exact(a, k, mode)) < tol
            assert O.rel_l2(O.oaconvolve(a, k, mode), O.conv_exact(a, k, mode)) < tol
    k = rng.standard_normal(95).astype(dt)
    for n in (200, 350, 500, 1000, 2000, 3000):
        a = rng.standard_normal(n).astype(dt)
        for mode in ("full", "same"):
            assert O.rel_l2(O.oaconvolve(a, k, mode), O.conv_exact(a, k, mode)) < tol


def test_same_mode_offset_is_floor_k_over_2_not_numpy():
    # include/fftconv/fftconv.hpp:369,415: differs from np.convolve(mode="same") for even K
    a = np.arange(1.0, 11.0)
    k = np.array([1.0, 2.0, 3.0, 4.0])
    full = np.convolve(a, k)
    np.testing.assert_allclose(O.convolve(a, k, "same"), full[2:12], atol=1e-12)
    np.testing.assert_allclose(O.oaconvolve(a, k, "same"), full[2:12], atol=1e-12)
    assert not np.allclose(np.convolve(a, k, "same"), full[2:12])


def test_python_binding_cases_on_oracle():
    # py/test/test_fftconv.py:9-70
    a = np.array([1, 2, 3, 4, 5], dtype=np.float64)
    k = np.array([0.2, 0.5, 0.2], dtype=np.float64)
    for fn in (O.convolve, O.oaconvolve):
        for mode in ("full", "same"):
            np.testing.assert_allclose(fn(a, k, mode), np.convolve(a, k, mode), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(O.hilbert(a), O.hilbert_exact(a), rtol=1e-5, atol=1e-8)


def test_envelope_restatement_against_scipy():
    """oracle.envelope / rf_envelope (post-processing of fftconv_cuda_envelope_*): the envelope
    underneath is the reference's hilbert; decimation and log compression follow the formula in
    include/fftconv_cuda.h, checked here against an independent float64 evaluation."""
    from scipy import signal

    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 1000))
    env = np.abs(signal.hilbert(x, axis=-1))
    got = O.envelope(x, log_compress=True, ref=2.0, dynamic_range_db=40.0, decimate=3)
    want = np.clip(1.0 + 20.0 * np.log10(env[:, ::3] / 2.0) / 40.0, 0.0, 1.0)
    assert got.shape == (4, 334)
    assert np.max(np.abs(got - want)) < 1e-12
    assert np.array_equal(O.envelope(x), O.hilbert(x))
    k = rng.standard_normal(65)
    filt = np.stack([np.convolve(r, k, "full")[32:32 + 1000] for r in x])
    assert O.rel_l2(O.rf_envelope(x, k), np.abs(signal.hilbert(filt, axis=-1))) < 1e-12
    z = np.zeros((1, 16))
    assert np.array_equal(O.envelope(z, log_compress=True), z)   # log of 0 clamps to 0
