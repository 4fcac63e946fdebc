"""ctypes loader for oracle/_ref/libfftconv_ref.so  (TEST INFRASTRUCTURE).

The library is the reference's own C++ headers (include/fftconv/*.hpp, unmodified,
compiled from /root/reference by oracle/Makefile) over oracle/fftw_shim.cpp.  It is
prebuilt in the build container and travels to the GPU box as a git-ignored .so;
nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libfftconv_ref.so")

_lib = None
_CT = {np.dtype(np.float32): ("f32", C.c_float), np.dtype(np.float64): ("f64", C.c_double)}


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                f"{LIB_PATH} missing: run `make -C oracle ref` in the build container "
                "(needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        L.ref_optimal_fft_size.restype = C.c_size_t
        L.ref_optimal_fft_size.argtypes = [C.c_size_t]
        L.ref_version.restype = C.c_char_p
        L.ref_hardware_threads.restype = C.c_int
        for sfx, ct in (("f32", C.c_float), ("f64", C.c_double)):
            P = C.POINTER(ct)
            for name in ("oaconvolve", "convolve"):
                f = getattr(L, f"ref_{name}_{sfx}")
                f.restype = None
                f.argtypes = [P, C.c_size_t, P, C.c_size_t, P, C.c_size_t, C.c_int]
                f = getattr(L, f"ref_{name}_batched_{sfx}")
                f.restype = None
                f.argtypes = [P, C.c_size_t, C.c_size_t, C.c_size_t, P, C.c_size_t, P,
                              C.c_size_t, C.c_size_t, C.c_int, C.c_int]
            f = getattr(L, f"ref_hilbert_{sfx}")
            f.restype = None
            f.argtypes = [P, P, C.c_size_t]
            f = getattr(L, f"ref_hilbert_batched_{sfx}")
            f.restype = None
            f.argtypes = [P, C.c_size_t, C.c_size_t, C.c_size_t, P, C.c_size_t, C.c_int]
            f = getattr(L, f"ref_rfft_{sfx}")
            f.restype = None
            f.argtypes = [P, C.c_size_t, P]
            f = getattr(L, f"ref_irfft_{sfx}")
            f.restype = None
            f.argtypes = [P, C.c_size_t, P]
        _lib = L
    return _lib


def _ptr(a):
    sfx, ct = _CT[a.dtype]
    return a.ctypes.data_as(C.POINTER(ct))


def _mode(mode) -> int:
    if mode in ("full", 0):
        return 0
    if mode in ("same", 1):
        return 1
    raise ValueError(f"Unsupported convolution mode: {mode}")


def _prep(a, k):
    a = np.ascontiguousarray(a)
    k = np.ascontiguousarray(k, dtype=a.dtype)
    if a.dtype not in _CT:
        raise TypeError("float32 or float64 only")
    return a, k


def optimal_fft_size(klen: int) -> int:
    return int(lib().ref_optimal_fft_size(klen))


def hardware_threads() -> int:
    return int(lib().ref_hardware_threads())


def _conv(name, a, k, mode, nthreads):
    a, k = _prep(a, k)
    m = _mode(mode)
    sfx = _CT[a.dtype][0]
    n = a.shape[-1]
    out_len = n + k.size - 1 if m == 0 else n
    if a.ndim == 1:
        out = np.empty(out_len, dtype=a.dtype)
        getattr(lib(), f"ref_{name}_{sfx}")(_ptr(a), n, _ptr(k), k.size, _ptr(out), out_len, m)
        return out
    if a.ndim != 2:
        raise ValueError("1-D or 2-D (batch, n) input only")
    out = np.empty((a.shape[0], out_len), dtype=a.dtype)
    getattr(lib(), f"ref_{name}_batched_{sfx}")(
        _ptr(a), a.shape[0], n, n, _ptr(k), k.size, _ptr(out), out_len, out_len, m, nthreads)
    return out


def oaconvolve(a, k, mode="full", nthreads=0):
    """reference oaconvolve_fftw (include/fftconv/fftconv.hpp:474-480) per row."""
    return _conv("oaconvolve", a, k, mode, nthreads)


def convolve(a, k, mode="full", nthreads=0):
    """reference convolve_fftw (include/fftconv/fftconv.hpp:453-461) per row."""
    return _conv("convolve", a, k, mode, nthreads)


def hilbert(x, nthreads=0):
    """reference hilbert (include/fftconv/hilbert.hpp:16-66) per row."""
    x = np.ascontiguousarray(x)
    sfx = _CT[x.dtype][0]
    out = np.empty_like(x)
    if x.ndim == 1:
        getattr(lib(), f"ref_hilbert_{sfx}")(_ptr(x), _ptr(out), x.size)
    else:
        n = x.shape[-1]
        getattr(lib(), f"ref_hilbert_batched_{sfx}")(_ptr(x), x.shape[0], n, n, _ptr(out), n, nthreads)
    return out


def rfft(x):
    x = np.ascontiguousarray(x)
    sfx = _CT[x.dtype][0]
    out = np.empty(2 * (x.size // 2 + 1), dtype=x.dtype)
    getattr(lib(), f"ref_rfft_{sfx}")(_ptr(x), x.size, _ptr(out))
    return out[0::2] + 1j * out[1::2]


def irfft_unnormalised(spec, n, dtype=np.float64):
    buf = np.empty(2 * (n // 2 + 1), dtype=dtype)
    buf[0::2] = spec.real
    buf[1::2] = spec.imag
    out = np.empty(n, dtype=dtype)
    sfx = _CT[np.dtype(dtype)][0]
    getattr(lib(), f"ref_irfft_{sfx}")(_ptr(buf), n, _ptr(out))
    return out
