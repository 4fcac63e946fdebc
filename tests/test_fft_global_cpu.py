"""CPU restatement of the index algebra of fftconv_b200/csrc/fft_global.cuh (the general FFT surface):
the Stockham pass schedule of gfft_pow2 and the chirp-z recipe of gfft_any, thread by thread in numpy,
against numpy.fft.  The kernels themselves are checked on the GPU (tests/test_gpu_round2.py)."""
import numpy as np
import pytest


def stockham_pass(src, n, ns, R, sign, W):
    """fc::gfft::stockham_pass<T, R> for one row."""
    dst = np.empty_like(src)
    per_row = n // R
    tstep = n // (ns * R)
    for j in range(per_row):
        k = j & (ns - 1)
        v = [src[j + r * per_row] for r in range(R)]
        for r in range(1, R):
            w = W[k * r * tstep]
            v[r] = v[r] * (np.conj(w) if sign > 0 else w)
        if R == 2:
            v = [v[0] + v[1], v[0] - v[1]]
        else:
            s0, d0, s1, d1 = v[0] + v[2], v[0] - v[2], v[1] + v[3], v[1] - v[3]
            jd1 = 1j * d1
            v = [s0 + s1, d0 - jd1, s0 - s1, d0 + jd1] if sign < 0 else [s0 + s1, d0 + jd1, s0 - s1, d0 - jd1]
        base = (j - k) * R + k
        for r in range(R):
            dst[base + r * ns] = v[r]
    return dst


def gfft_pow2(x, sign):
    m = len(x)
    if m == 1:
        return x.copy()
    W = np.exp(-2j * np.pi * np.arange(m) / m)
    l2 = m.bit_length() - 1
    passes = l2 // 2 + (l2 & 1)
    ns, src = 1, x
    for p in range(1, passes + 1):
        R = 2 if (l2 & 1) and p == passes else 4
        src = stockham_pass(src, m, ns, R, sign, W)
        ns *= R
    return src


def gfft_any(x, sign):
    n = len(x)
    if n & (n - 1) == 0:
        return gfft_pow2(x, sign)
    M = 1
    while M < 2 * n - 1:
        M *= 2
    j = np.arange(n)
    chirp = np.exp(-1j * np.pi * ((j * j) % (2 * n)) / n)
    b = np.zeros(M, complex)
    b[:n] = np.conj(chirp)
    b[M - n + 1:] = np.conj(chirp[1:][::-1])
    Bf = gfft_pow2(b, -1)
    cj = sign > 0
    a = np.zeros(M, complex)
    a[:n] = (np.conj(x) if cj else x) * chirp
    y = gfft_pow2(gfft_pow2(a, -1) * Bf, +1)
    out = y[:n] * chirp / M
    return np.conj(out) if cj else out


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 512, 3, 5, 6, 7, 12, 100, 243, 1000])
def test_stockham_and_chirpz_index_algebra(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    np.testing.assert_allclose(gfft_any(x, -1), np.fft.fft(x), rtol=0, atol=1e-9 * max(1, n))
    np.testing.assert_allclose(gfft_any(x, +1), np.fft.ifft(x) * n, rtol=0, atol=1e-9 * max(1, n))
