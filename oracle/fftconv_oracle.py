"""NumPy restatement of the reference convolution hot path -- TEST INFRASTRUCTURE.

Parity status: PINNED.  This restatement is checked (tests/test_oracle.py) against
  * the reference's literal known-answer vectors: the 10-sample Hilbert vector
    (test/test_hilbert.cpp:8-19) and the 20-point DFT vectors that pin FFTW's sign
    and scaling conventions (test/test_fftw.cpp:261-299, 347-393);
  * the reference's own headers compiled unmodified over oracle/fftw_shim.cpp
    (oracle/_ref, outputs committed as tests/golden/*.npz by tests/golden/make_golden.py);
  * the mathematical definition (float64 direct convolution), which is what the
    reference's own tests use as oracle (test/test_fftconv.cpp:93-218 via arma::conv).

The arithmetic of the reference lives in FFTW 3.3.10 (third party, un-vendored,
vcpkg.json:5-23); here numpy.fft (pocketfft) plays that role, so agreement between
this file and oracle/_ref is agreement between two independent FFTs.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import numpy as np

# include/fftconv/fftconv.hpp:38-49 -- {max_filter_size, optimal_fft_size}; the CSV
# optimal_fft_size.csv:1-9 holds the same rows.
OPTIMAL_OA_FFT_SIZE = (
    (7, 16), (12, 32), (21, 64), (36, 128), (65, 256),
    (120, 512), (221, 1024), (411, 2048), (769, 4096),
)
MAX_OACONV_FFT_SIZE = 8192  # include/fftconv/fftconv.hpp:49


def get_optimal_fft_size(filter_size: int) -> int:
    """include/fftconv/fftconv.hpp:53-66."""
    for max_filter, fft_size in OPTIMAL_OA_FFT_SIZE:
        if filter_size < max_filter:
            return fft_size
    size = MAX_OACONV_FFT_SIZE
    while size < filter_size * 2:
        size *= 2
    return size


def cost_model_table(max_m: int = 1000) -> list[tuple[int, int]]:
    """Re-derivation of optimal_fft_size.csv from cost_func.py:5-13,25-34.

    cost(N, M) = N (log2 N + 1) / max(N - M + 1, 0) over N = 2^4..2^13, M = 1..999,
    truncated to an integer, first arg-min wins.  The committed CSV is reproduced
    only when an infeasible N (N - M + 1 <= 0, cost = +inf) stays +infinity after
    the integer cast (saturating conversion, as on the author's ARM machine); on
    x86 NumPy casts inf to INT64_MIN, which would win every arg-min (SURVEY.md
    section 8a row 1).
    """
    ms = np.arange(1, max_m, dtype=np.int64)
    ns = 2 ** np.arange(4, 14, dtype=np.int64)
    nn, mm = np.meshgrid(ns, ms)
    denom = np.maximum(nn - mm + 1, 0).astype(np.float64)
    with np.errstate(divide="ignore"):
        cost = nn * (np.log2(nn) + 1) / denom
    icost = np.where(np.isfinite(cost), np.floor(cost), np.inf)  # saturating int cast
    best = ns[np.argmin(icost, axis=1)]
    pairs, last = [], best[0]
    for m, n in zip(ms, best):
        if n == last:
            continue
        pairs.append((int(m), int(last)))
        last = n
    return pairs


def _cdtype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


def _check(a, k):
    a = np.asarray(a)
    k = np.asarray(k, dtype=a.dtype)
    if a.dtype not in (np.float32, np.float64):
        raise TypeError("float32 or float64 only (fftw::Floating, fftw.hpp:47-48)")
    return a, k


def oaconvolve(a, k, mode: str = "full"):
    """FFTConvEngine::oaconvolve, include/fftconv/fftconv.hpp:382-437.

    ``a`` is (n,) or (batch, n); every row is convolved with the same taps ``k``.
    Arithmetic stays in the input precision (complex64 spectra for float32), as in
    the reference.  Segments of a row are processed in order and accumulated into
    ``out`` exactly like the reference's ``normalize_add`` (fftw.hpp:627-643).
    """
    a, k = _check(a, k)
    squeeze = a.ndim == 1
    a2 = np.atleast_2d(a)
    n, klen = a2.shape[-1], k.size
    nfft = get_optimal_fft_size(klen)
    step = nfft - (klen - 1)                      # :387-388
    cd = _cdtype(a.dtype)
    h = np.fft.rfft(k, nfft).astype(cd)           # :391-392, kernel spectrum each call
    # :394 fct = 1/N is applied inside numpy's irfft
    if mode == "full":
        out = np.zeros((a2.shape[0], n + klen - 1), dtype=a.dtype)   # :385, :397
    elif mode == "same":
        out = np.zeros((a2.shape[0], n), dtype=a.dtype)               # :413
    else:
        raise ValueError(f"Unsupported convolution mode: {mode}")    # py/main.cpp:21-27
    pad = klen // 2                               # :415
    for pos in range(0, n, step):                 # :400 / :417
        seg = a2[:, pos:pos + step]               # len = min(n - pos, step)
        spec = np.fft.rfft(seg, nfft, axis=-1).astype(cd)            # :403-404
        y = np.fft.irfft(spec * h, nfft, axis=-1).astype(a.dtype)    # :405-406 (irfft scales by 1/N)
        if mode == "full":
            m = min(nfft, out.shape[1] - pos)     # normalize_add min(), fftw.hpp:639
            out[:, pos:pos + m] += y[:, :m]
        elif pos < pad:                           # :426-428
            m = min(out.shape[1], nfft - (pad - pos))
            out[:, :m] += y[:, pad - pos:pad - pos + m]
        else:                                     # :429-431
            m = min(nfft, out.shape[1] - (pos - pad))
            out[:, pos - pad:pos - pad + m] += y[:, :m]
    return out[0] if squeeze else out


def convolve(a, k, mode: str = "full"):
    """FFTConvEngine::convolve, include/fftconv/fftconv.hpp:338-380 (one FFT of
    length n + K - 1, any length)."""
    a, k = _check(a, k)
    squeeze = a.ndim == 1
    a2 = np.atleast_2d(a)
    n, klen = a2.shape[-1], k.size
    padded = n + klen - 1                         # :458
    cd = _cdtype(a.dtype)
    sa = np.fft.rfft(a2, padded, axis=-1).astype(cd)      # :344-345
    sk = np.fft.rfft(k, padded).astype(cd)                # :349-350
    y = np.fft.irfft(sa * sk, padded, axis=-1).astype(a.dtype)   # :353-365
    if mode == "full":
        out = y
    elif mode == "same":
        pad = klen // 2                           # :369
        out = y[:, pad:pad + n]                   # :377-378
    else:
        raise ValueError(f"Unsupported convolution mode: {mode}")
    out = np.ascontiguousarray(out)
    return out[0] if squeeze else out


def hilbert(x):
    """fftconv::hilbert, include/fftconv/hilbert.hpp:16-66: envelope |x + j H{x}|."""
    x = np.asarray(x)
    if x.dtype not in (np.float32, np.float64):
        raise TypeError("float32 or float64 only")
    n = x.shape[-1]
    cd = _cdtype(x.dtype)
    spec = np.fft.rfft(x, axis=-1).astype(cd)             # :25-36
    rot = np.empty_like(spec)
    rot.real = spec.imag                                  # :46-49  (re, im) -> (im, -re)
    rot.imag = -spec.real
    rot[..., 0] = 0                                       # :42-44  DC
    if n % 2 == 0:
        rot[..., -1] = 0                                  # :42-44  Nyquist (n even)
    h = np.fft.irfft(rot, n, axis=-1).astype(x.dtype)     # :54, :57 (1/n inside irfft)
    return np.sqrt(x * x + h * h).astype(x.dtype)         # :59-63


def envelope(x, log_compress: bool = False, ref: float = 1.0, dynamic_range_db: float = 60.0, decimate: int = 1):
    """Envelope post-processing (include/fftconv_cuda.h, fftconv_cuda_envelope_*; SURVEY.md section
    8f item 3) restated on top of the reference's hilbert (include/fftconv/hilbert.hpp:16-66):
    env = hilbert(x); y[j] = env[j * decimate]; with log_compress
    y = clamp(1 + 20 log10(y / ref) / dynamic_range_db, 0, 1).  The reference has no such
    function: the definition is this one, the envelope underneath is the reference's."""
    env = hilbert(x)
    y = env[..., ::int(decimate)]
    if log_compress:
        with np.errstate(divide="ignore"):
            db = 20.0 * np.log10(y.astype(np.float64) / float(ref))
        y = np.clip(1.0 + db / float(dynamic_range_db), 0.0, 1.0).astype(env.dtype)
    return np.ascontiguousarray(y)


def rf_envelope(x, k, log_compress: bool = False, ref: float = 1.0, dynamic_range_db: float = 60.0,
                decimate: int = 1):
    """oaconvolve_fftw<T, Same> (include/fftconv/fftconv.hpp:474-480) row by row, then `envelope`."""
    return envelope(oaconvolve(x, k, "same"), log_compress, ref, dynamic_range_db, decimate)


# ---------------------------------------------------------------------------
# float64 ground truth by the mathematical definition (the reference tests'
# oracle is arma::conv = direct convolution, test/test_fftconv.cpp:100,116,137).
# ---------------------------------------------------------------------------
def conv_exact(a, k, mode: str = "full"):
    """Direct convolution in float64; Same starts at floor(K/2) of the full result
    (include/fftconv/fftconv.hpp:369,415) -- NOT numpy's floor((K-1)/2)."""
    a = np.asarray(a, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    rows = np.atleast_2d(a)
    full = np.stack([np.convolve(r, k, "full") for r in rows])
    if mode == "same":
        pad = k.size // 2
        full = full[:, pad:pad + rows.shape[1]]
    elif mode != "full":
        raise ValueError(mode)
    return full[0] if a.ndim == 1 else full


def conv_exact_at(a_row, k, idx):
    """Full-mode outputs at selected indices only (float64 dot products); used for
    sampled checks of signals too long to convolve densely."""
    a_row = np.asarray(a_row, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    n, klen = a_row.size, k.size
    out = np.empty(len(idx), dtype=np.float64)
    for t, i in enumerate(idx):
        lo, hi = max(0, i - klen + 1), min(n - 1, i)
        seg = a_row[lo:hi + 1]
        taps = k[i - hi:i - lo + 1][::-1]
        out[t] = float(np.dot(seg, taps))
    return out


def hilbert_exact(x):
    from scipy import signal
    return np.abs(signal.hilbert(np.asarray(x, dtype=np.float64), axis=-1))


def rel_l2(y, ref) -> float:
    y = np.asarray(y, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    d = np.linalg.norm(ref)
    return float(np.linalg.norm(y - ref) / (d if d > 0 else 1.0))
