"""ctypes binding of libfftconv_cuda.so (include/fftconv_cuda.h).

The product path: there is no CPU fallback.  If the shared library is missing this
module raises at import of the first symbol; if no CUDA device is usable every call
raises RuntimeError with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFTCONV_B200_LIB: another build of the same library (A/B measurements of two builds on one box)
LIB_PATH = os.environ.get("FFTCONV_B200_LIB") or os.path.join(_HERE, "lib", "libfftconv_cuda.so")

FULL, SAME = 0, 1

# every symbol include/fftconv_cuda.h declares
SYMBOLS = [
    "fftconv_cuda_version", "fftconv_cuda_last_error", "fftconv_cuda_optimal_fft_size",
    "fftconv_cuda_clear_cache", "fftconv_cuda_launch_count", "fftconv_cuda_set_reserved_sms",
    "fftconv_cuda_release_stream",
    "fftconv_cuda_mgpu_init", "fftconv_cuda_mgpu_destroy", "fftconv_cuda_mgpu_device_count",
    "fftconv_cuda_mgpu_sync", "fftconv_cuda_mgpu_chunk_bounds", "fftconv_cuda_mgpu_row_bounds",
] + [
    f"fftconv_cuda_{name}_{sfx}"
    for sfx in ("f32", "f64")
    for name in ("oaconvolve", "oaconvolve_batched", "convolve", "convolve_batched",
                 "hilbert", "hilbert_batched", "add",
                 "filter_create", "filter_oaconvolve", "stream_create", "stream_push",
                 "stream_flush", "envelope_batched", "rf_envelope_batched",
                 "oa_tail", "fill_synthetic",
                 "mgpu_oaconvolve_batched", "mgpu_hilbert_batched", "mgpu_oaconvolve_long",
                 "mgpu_oaconvolve_long_host", "fft_c2c", "fft_r2c", "fft_c2r")
] + ["fftconv_cuda_filter_destroy", "fftconv_cuda_stream_destroy"]



class EnvelopeOpts(C.Structure):
    """fftconv_cuda_envelope_opts (include/fftconv_cuda.h)."""
    _fields_ = [("log_compress", C.c_int), ("ref", C.c_double), ("dynamic_range_db", C.c_double),
                ("decimate", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m fftconv_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.fftconv_cuda_version.restype = C.c_char_p
    L.fftconv_cuda_last_error.restype = C.c_char_p
    L.fftconv_cuda_optimal_fft_size.restype = C.c_size_t
    L.fftconv_cuda_optimal_fft_size.argtypes = [C.c_size_t]
    L.fftconv_cuda_clear_cache.restype = C.c_int
    L.fftconv_cuda_launch_count.restype = C.c_ulonglong
    L.fftconv_cuda_set_reserved_sms.restype, L.fftconv_cuda_set_reserved_sms.argtypes = C.c_int, [C.c_int]
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    L.fftconv_cuda_release_stream.restype, L.fftconv_cuda_release_stream.argtypes = i, [vp]
    L.fftconv_cuda_mgpu_init.restype, L.fftconv_cuda_mgpu_init.argtypes = i, [i, C.POINTER(i), C.POINTER(vp)]
    L.fftconv_cuda_mgpu_destroy.restype, L.fftconv_cuda_mgpu_destroy.argtypes = i, [vp]
    L.fftconv_cuda_mgpu_device_count.restype, L.fftconv_cuda_mgpu_device_count.argtypes = i, [vp]
    L.fftconv_cuda_mgpu_sync.restype, L.fftconv_cuda_mgpu_sync.argtypes = i, [vp]
    for name in ("chunk_bounds", "row_bounds"):
        f = getattr(L, f"fftconv_cuda_mgpu_{name}")
        f.restype, f.argtypes = None, [sz, i, i, C.POINTER(sz), C.POINTER(sz)]
    for sfx in ("f32", "f64"):
        f = getattr(L, f"fftconv_cuda_oa_tail_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, vp, vp]
        f = getattr(L, f"fftconv_cuda_fill_synthetic_{sfx}")
        f.restype, f.argtypes = i, [vp, sz, C.c_ulonglong, C.c_ulonglong, C.c_double, vp]
        f = getattr(L, f"fftconv_cuda_fft_c2c_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, sz, i, vp]
        for name in ("r2c", "c2r"):
            f = getattr(L, f"fftconv_cuda_fft_{name}_{sfx}")
            f.restype, f.argtypes = i, [vp, vp, sz, sz, vp]
        f = getattr(L, f"fftconv_cuda_mgpu_oaconvolve_batched_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, sz, sz, vp, sz, vp, sz, sz, i]
        f = getattr(L, f"fftconv_cuda_mgpu_hilbert_batched_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, sz, sz, vp, sz]
        f = getattr(L, f"fftconv_cuda_mgpu_oaconvolve_long_{sfx}")
        f.restype, f.argtypes = i, [vp, C.POINTER(vp), C.POINTER(sz), vp, sz, C.POINTER(vp)]
        f = getattr(L, f"fftconv_cuda_mgpu_oaconvolve_long_host_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, vp, sz, vp]
    for sfx in ("f32", "f64"):
        for name in ("oaconvolve", "convolve"):
            f = getattr(L, f"fftconv_cuda_{name}_{sfx}")
            f.restype, f.argtypes = i, [vp, sz, vp, sz, vp, sz, i]
            f = getattr(L, f"fftconv_cuda_{name}_batched_{sfx}")
            f.restype, f.argtypes = i, [vp, sz, sz, sz, vp, sz, vp, sz, sz, i, vp]
        f = getattr(L, f"fftconv_cuda_hilbert_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz]
        f = getattr(L, f"fftconv_cuda_hilbert_batched_{sfx}")
        f.restype, f.argtypes = i, [vp, sz, sz, sz, vp, sz, vp]
        f = getattr(L, f"fftconv_cuda_add_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, vp]
        f = getattr(L, f"fftconv_cuda_filter_create_{sfx}")
        f.restype, f.argtypes = i, [vp, sz, C.POINTER(vp)]
        f = getattr(L, f"fftconv_cuda_filter_oaconvolve_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, sz, sz, vp, sz, sz, i, vp]
        f = getattr(L, f"fftconv_cuda_stream_create_{sfx}")
        f.restype, f.argtypes = i, [vp, sz, C.POINTER(vp)]
        f = getattr(L, f"fftconv_cuda_stream_push_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, sz, vp, sz, vp]
        f = getattr(L, f"fftconv_cuda_stream_flush_{sfx}")
        f.restype, f.argtypes = i, [vp, vp, sz, vp]
        f = getattr(L, f"fftconv_cuda_envelope_batched_{sfx}")
        f.restype, f.argtypes = i, [vp, sz, sz, sz, vp, sz, sz, C.POINTER(EnvelopeOpts), vp]
        f = getattr(L, f"fftconv_cuda_rf_envelope_batched_{sfx}")
        f.restype, f.argtypes = i, [vp, sz, sz, sz, vp, sz, vp, sz, sz, C.POINTER(EnvelopeOpts), vp]
    L.fftconv_cuda_filter_destroy.restype, L.fftconv_cuda_filter_destroy.argtypes = i, [vp]
    L.fftconv_cuda_stream_destroy.restype, L.fftconv_cuda_stream_destroy.argtypes = i, [vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().fftconv_cuda_last_error().decode("utf-8", "replace")
        raise RuntimeError(msg or f"libfftconv_cuda error {rc}")
