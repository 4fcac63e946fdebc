"""Python face of the hot path, mirroring the reference's `pyfftconv` module
(/root/reference/py/main.cpp:147-174, stubs py/pyfftconv/_pyfftconv.pyi):

    convolve(a, k, mode="full")      convolve_(a, k, out, mode="full")
    oaconvolve(a, k, mode="full")    oaconvolve_(a, k, out, mode="full")
    hilbert(a)                       hilbert_(a, out)

Same names, argument meaning, defaults and error behaviour (RuntimeError for a
kernel longer than the signal or an unknown mode).  Extensions over the reference:
2-D inputs are batches of rows convolved with the same taps (the reference's
binding rejects ndim != 1), and CUDA `torch.Tensor`s -- or, in the `_` forms, any object with
`__cuda_array_interface__` (CuPy, Numba) or `__dlpack__` on a CUDA device -- are processed in
place on the device on the caller's stream (no host round trip).

Everything goes through the C ABI of libfftconv_cuda.so; nothing here computes.
"""
from __future__ import annotations

import numpy as np

from . import _capi

try:  # torch is plumbing (device memory, streams); numpy-only callers do not need it
    import torch
except Exception:  # pragma: no cover
    torch = None

_SFX = {"float32": "f32", "float64": "f64"}


def _mode(mode) -> int:
    # py/main.cpp:21-27
    if mode == "same":
        return _capi.SAME
    if mode == "full":
        return _capi.FULL
    raise RuntimeError(f"Unsupported convolution mode: {mode}")


def _is_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class _Buf:
    """pointer + shape view of a numpy array or a torch tensor (rows contiguous)."""

    def __init__(self, x, name: str):
        if (torch is not None and not _is_tensor(x) and not isinstance(x, np.ndarray)
                and not hasattr(x, "__cuda_array_interface__") and hasattr(x, "__dlpack__")
                and hasattr(x, "__dlpack_device__") and int(x.__dlpack_device__()[0]) == 2):
            x = torch.from_dlpack(x)     # any DLPack exporter on a CUDA device: zero-copy view
        self.obj = x
        if not _is_tensor(x) and hasattr(x, "__cuda_array_interface__"):
            # CuPy / Numba / any CUDA array: taken in place, no torch needed
            ci = x.__cuda_array_interface__
            self.dtype = {"<f4": "float32", "<f8": "float64"}.get(ci["typestr"], ci["typestr"])
            self.shape = tuple(ci["shape"])
            item = 4 if self.dtype == "float32" else 8
            strides = ci.get("strides")
            if strides is None:
                strides = tuple(item * int(np.prod(self.shape[i + 1:])) for i in range(len(self.shape)))
            if len(self.shape) >= 1 and self.shape[-1] > 1 and strides[-1] != item:
                raise RuntimeError(f"{name}: last dimension must be contiguous")
            self.row_stride = strides[0] // item if len(self.shape) == 2 else (self.shape[-1] if self.shape else 1)
            self.ptr = int(ci["data"][0])
            self.is_cuda = True
            self.cai_stream = ci.get("stream")
        elif _is_tensor(x):
            self.dtype = str(x.dtype).replace("torch.", "")
            self.shape = tuple(x.shape)
            if x.dim() >= 1 and x.stride(-1) != 1 and x.shape[-1] > 1:
                raise RuntimeError(f"{name}: last dimension must be contiguous")
            self.row_stride = x.stride(0) if x.dim() == 2 else (x.shape[-1] if x.dim() else 1)
            self.ptr = x.data_ptr()
            self.is_cuda = x.is_cuda
        else:
            x = np.asarray(x)
            self.obj = x
            self.dtype = x.dtype.name                     # (str(dtype) costs ~6 us)
            self.shape = x.shape
            if x.ndim >= 1 and x.shape[-1] > 1 and x.strides[-1] != x.itemsize:
                raise RuntimeError(f"{name}: last dimension must be contiguous")
            self.row_stride = (x.strides[0] // x.itemsize) if x.ndim == 2 else (x.shape[-1] if x.ndim else 1)
            self.ptr = x.__array_interface__["data"][0]   # (x.ctypes.data costs ~3 us per array)
            self.is_cuda = False
        if self.dtype not in _SFX:
            raise TypeError(f"{name}: float32 or float64 expected, got {self.dtype}")
        self.ndim = len(self.shape)


def _stream_of(buf: _Buf):
    if not buf.is_cuda:
        return None
    if _is_tensor(buf.obj):
        return torch.cuda.current_stream(buf.obj.device).cuda_stream
    s = getattr(buf, "cai_stream", None)     # __cuda_array_interface__ v3: None / 1 = legacy default stream,
    if s == 2:                                # 2 = per-thread default stream (cudaStreamPerThread, handle 0x2)
        return 2
    return None if s in (None, 1) else int(s)


class _device_of:
    """Context: make the array's device current (torch tensors know theirs; other CUDA
    arrays are taken to live on the current device, as their own libraries assume)."""

    def __init__(self, buf: _Buf):
        # switching devices costs several microseconds per call: only when it is not current already
        self.ctx = None
        if _is_tensor(buf.obj) and buf.obj.device.index not in (None, torch.cuda.current_device()):
            self.ctx = torch.cuda.device(buf.obj.device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _empty_like_rows(a, out_len: int):
    if _is_tensor(a):
        shape = (out_len,) if a.dim() == 1 else (a.shape[0], out_len)
        return torch.empty(shape, dtype=a.dtype, device=a.device)
    a = np.asarray(a)
    shape = (out_len,) if a.ndim == 1 else (a.shape[0], out_len)
    return np.empty(shape, dtype=a.dtype)


def _conv_(name: str, a, k, out, mode):
    A, K, O = _Buf(a, "a"), _Buf(k, "k"), _Buf(out, "out")
    # py/main.cpp:38-44 / :85-91
    if K.ndim != 1 or A.ndim not in (1, 2) or O.ndim != A.ndim:
        raise RuntimeError("Number of dimensions must be one")
    if A.dtype != K.dtype or A.dtype != O.dtype:
        raise TypeError("a, k and out must share one dtype")
    n, klen = A.shape[-1], K.shape[0]
    if n < klen:
        raise RuntimeError("Kernel size must be smaller than input size")
    m = _mode(mode)
    out_len = O.shape[-1]
    batch = 1 if A.ndim == 1 else A.shape[0]
    if A.ndim == 2 and O.shape[0] != batch:
        raise RuntimeError("out must have one row per input row")
    if A.is_cuda != O.is_cuda:
        raise RuntimeError("a and out must live on the same side (host or device)")
    fn = getattr(_capi.lib(), f"fftconv_cuda_{name}_batched_{_SFX[A.dtype]}")
    if A.is_cuda:
        with _device_of(A):
            rc = fn(A.ptr, batch, n, A.row_stride, K.ptr, klen, O.ptr, out_len, O.row_stride, m,
                    _stream_of(A))
    else:
        rc = fn(A.ptr, batch, n, A.row_stride, K.ptr, klen, O.ptr, out_len, O.row_stride, m, None)
    _capi.check(rc)


def _conv(name: str, a, k, mode):
    m = _mode(mode)
    n = a.shape[-1]
    klen = k.shape[0] if hasattr(k, "shape") else len(k)
    out = _empty_like_rows(a, n if m == _capi.SAME else n + klen - 1)  # py/main.cpp:62-71
    _conv_(name, a, k, out, mode)
    return out


def convolve_(a, k, out, mode: str = "full") -> None:
    """fftconv::convolve_fftw into `out` (py/main.cpp:29-55)."""
    _conv_("convolve", a, k, out, mode)


def convolve(a, k, mode: str = "full"):
    """fftconv::convolve_fftw; API compatible with np.convolve (py/main.cpp:58-74)."""
    return _conv("convolve", a, k, mode)


def oaconvolve_(a, k, out, mode: str = "full") -> None:
    """fftconv::oaconvolve_fftw into `out` (py/main.cpp:76-102)."""
    _conv_("oaconvolve", a, k, out, mode)


def oaconvolve(a, k, mode: str = "full"):
    """fftconv::oaconvolve_fftw (py/main.cpp:105-121)."""
    return _conv("oaconvolve", a, k, mode)


def hilbert_(a, out) -> None:
    """fftconv::hilbert into `out`: envelope |a + j H{a}| (py/main.cpp:123-126)."""
    A, O = _Buf(a, "a"), _Buf(out, "out")
    if A.ndim not in (1, 2) or O.ndim != A.ndim or A.shape != O.shape:
        raise RuntimeError("hilbert: a and out must have the same 1-D or 2-D shape")
    if A.dtype != O.dtype:
        raise TypeError("a and out must share one dtype")
    if A.is_cuda != O.is_cuda:
        raise RuntimeError("a and out must live on the same side (host or device)")
    n = A.shape[-1]
    batch = 1 if A.ndim == 1 else A.shape[0]
    fn = getattr(_capi.lib(), f"fftconv_cuda_hilbert_batched_{_SFX[A.dtype]}")
    if A.is_cuda:
        with _device_of(A):
            rc = fn(A.ptr, batch, n, A.row_stride, O.ptr, O.row_stride, _stream_of(A))
    else:
        rc = fn(A.ptr, batch, n, A.row_stride, O.ptr, O.row_stride, None)
    _capi.check(rc)


def hilbert(a):
    """Equivalent to np.abs(scipy.signal.hilbert(a)) (py/main.cpp:128-132, :142-145)."""
    out = _empty_like_rows(a, a.shape[-1])
    hilbert_(a, out)
    return out


def _envelope_opts(log_compress, ref, dynamic_range_db, decimate):
    decimate = int(decimate)
    if decimate < 1:
        raise RuntimeError("decimate must be >= 1")
    return _capi.EnvelopeOpts(1 if log_compress else 0, float(ref), float(dynamic_range_db), decimate)


def envelope_out_len(n: int, decimate: int = 1) -> int:
    return (int(n) + int(decimate) - 1) // int(decimate)


def _envelope_(a, k, out, log_compress, ref, dynamic_range_db, decimate) -> None:
    A, O = _Buf(a, "a"), _Buf(out, "out")
    opts = _envelope_opts(log_compress, ref, dynamic_range_db, decimate)
    if A.ndim not in (1, 2) or O.ndim != A.ndim:
        raise RuntimeError("envelope: a and out must both be 1-D or both 2-D")
    n = A.shape[-1]
    batch = 1 if A.ndim == 1 else A.shape[0]
    want = envelope_out_len(n, opts.decimate)
    if O.shape[-1] != want or (A.ndim == 2 and O.shape[0] != batch):
        raise RuntimeError(f"envelope: out must hold {want} samples per row (ceil(n / decimate))")
    if A.dtype != O.dtype:
        raise TypeError("a and out must share one dtype")
    if A.is_cuda != O.is_cuda:
        raise RuntimeError("a and out must live on the same side (host or device)")
    stream = None
    sfx = _SFX[A.dtype]
    if k is None:
        fn = getattr(_capi.lib(), f"fftconv_cuda_envelope_batched_{sfx}")
        args = (A.ptr, batch, n, A.row_stride, O.ptr, want, O.row_stride)
    else:
        K = _Buf(k, "k")
        if K.ndim != 1 or K.dtype != A.dtype:
            raise RuntimeError("rf_envelope: k must be 1-D and share a's dtype")
        if n < K.shape[0]:
            raise RuntimeError("a should be no shorter than k")  # py/main.cpp:88-91
        fn = getattr(_capi.lib(), f"fftconv_cuda_rf_envelope_batched_{sfx}")
        args = (A.ptr, batch, n, A.row_stride, K.ptr, K.shape[0], O.ptr, want, O.row_stride)
    if A.is_cuda:
        with _device_of(A):
            rc = fn(*args, opts, _stream_of(A))
    else:
        rc = fn(*args, opts, stream)
    _capi.check(rc)


def envelope_(a, out, *, log_compress: bool = False, ref: float = 1.0, dynamic_range_db: float = 60.0,
              decimate: int = 1) -> None:
    """`hilbert` with the post-processing its callers apply next fused into the kernel's store
    (SURVEY.md section 8f item 3): keep every `decimate`-th envelope sample, and with
    `log_compress` map it to clamp(1 + 20 log10(env / ref) / dynamic_range_db, 0, 1)."""
    _envelope_(a, None, out, log_compress, ref, dynamic_range_db, decimate)


def envelope(a, *, log_compress: bool = False, ref: float = 1.0, dynamic_range_db: float = 60.0,
             decimate: int = 1):
    out = _empty_like_rows(a, envelope_out_len(a.shape[-1], decimate))
    envelope_(a, out, log_compress=log_compress, ref=ref, dynamic_range_db=dynamic_range_db, decimate=decimate)
    return out


def rf_envelope_(a, k, out, *, log_compress: bool = False, ref: float = 1.0, dynamic_range_db: float = 60.0,
                 decimate: int = 1) -> None:
    """Band-pass `oaconvolve(a, k, "same")`, then `envelope` of the filtered rows; the filtered
    rows stay on the device between the two kernels."""
    _envelope_(a, k, out, log_compress, ref, dynamic_range_db, decimate)


def rf_envelope(a, k, *, log_compress: bool = False, ref: float = 1.0, dynamic_range_db: float = 60.0,
                decimate: int = 1):
    out = _empty_like_rows(a, envelope_out_len(a.shape[-1], decimate))
    rf_envelope_(a, k, out, log_compress=log_compress, ref=ref, dynamic_range_db=dynamic_range_db,
                 decimate=decimate)
    return out


# ---- general 1-D FFT surface (fftconv_cuda_fft_*) ------------------------
_CPLX = {"complex64": ("f32", "float32"), "complex128": ("f64", "float64")}
_REAL = {"float32": ("f32", "complex64"), "float64": ("f64", "complex128")}


def _fft_call(name: str, x, out, n: int, batch: int, sign=None) -> None:
    """x / out: contiguous numpy arrays or CUDA torch tensors (both of one kind)."""
    if _is_tensor(x):
        if not (x.is_cuda and out.is_cuda):
            raise RuntimeError("fft: torch tensors must be CUDA tensors (host data: numpy arrays)")
        xin = torch.view_as_real(x) if x.is_complex() else x
        xout = torch.view_as_real(out) if out.is_complex() else out
        sfx = "f32" if xin.dtype == torch.float32 else "f64"
        fn = getattr(_capi.lib(), f"fftconv_cuda_fft_{name}_{sfx}")
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream(x.device).cuda_stream
            args = (xin.data_ptr(), xout.data_ptr(), n, batch) + ((sign,) if sign is not None else ()) + (st,)
            _capi.check(fn(*args))
        return
    sfx = "f32" if x.dtype in (np.float32, np.complex64) else "f64"
    fn = getattr(_capi.lib(), f"fftconv_cuda_fft_{name}_{sfx}")
    args = (x.ctypes.data, out.ctypes.data, n, batch) + ((sign,) if sign is not None else ()) + (None,)
    _capi.check(fn(*args))


def _rows(x, name):
    if _is_tensor(x):
        if x.dim() not in (1, 2):
            raise RuntimeError(f"{name}: 1-D or 2-D input expected")
        return x.contiguous()
    x = np.ascontiguousarray(x)
    if x.ndim not in (1, 2):
        raise RuntimeError(f"{name}: 1-D or 2-D input expected")
    return x


def _new(x, shape, dtype_name):
    if _is_tensor(x):
        return torch.empty(shape, dtype=getattr(torch, dtype_name), device=x.device)
    return np.empty(shape, dtype=dtype_name)


def _dtype_name(x) -> str:
    return str(x.dtype).replace("torch.", "") if _is_tensor(x) else x.dtype.name


def fft(x, *, inverse: bool = False):
    """Complex transform along the last axis of a 1-D or 2-D complex64 / complex128 array (any length):
    numpy.fft.fft; with inverse=True FFTW's BACKWARD transform, i.e. numpy.fft.ifft(x) * n
    (the reference's fftw::Plan<T>::dft_1d + execute, include/fftconv/fftw.hpp:152-155)."""
    x = _rows(x, "fft")
    dn = _dtype_name(x)
    if dn not in _CPLX:
        raise TypeError(f"fft: complex64 or complex128 expected, got {dn}")
    out = _new(x, tuple(x.shape), dn)
    if x.shape[-1] == 0:
        raise RuntimeError("fft: empty transform")
    _fft_call("c2c", x, out, x.shape[-1], 1 if len(x.shape) == 1 else x.shape[0], +1 if inverse else -1)
    return out


def rfft(x):
    """numpy.fft.rfft along the last axis of a real float32 / float64 array: n // 2 + 1 points per row
    (fftw::Plan<T>::dft_r2c_1d, include/fftconv/fftw.hpp:174-177)."""
    x = _rows(x, "rfft")
    dn = _dtype_name(x)
    if dn not in _REAL:
        raise TypeError(f"rfft: float32 or float64 expected, got {dn}")
    n = x.shape[-1]
    if n == 0:
        raise RuntimeError("rfft: empty transform")
    out = _new(x, tuple(x.shape[:-1]) + (n // 2 + 1,), _REAL[dn][1])
    _fft_call("r2c", x, out, n, 1 if len(x.shape) == 1 else x.shape[0])
    return out


def irfft(x, n: int):
    """FFTW's c2r transform (unnormalised: numpy.fft.irfft(x, n) * n) of rows of n // 2 + 1 complex points
    (fftw::Plan<T>::dft_c2r_1d, include/fftconv/fftw.hpp:191-194)."""
    x = _rows(x, "irfft")
    dn = _dtype_name(x)
    if dn not in _CPLX:
        raise TypeError(f"irfft: complex64 or complex128 expected, got {dn}")
    if n <= 0 or x.shape[-1] != n // 2 + 1:
        raise RuntimeError("irfft: rows must hold n // 2 + 1 points")
    out = _new(x, tuple(x.shape[:-1]) + (n,), _CPLX[dn][1])
    _fft_call("c2r", x, out, n, 1 if len(x.shape) == 1 else x.shape[0])
    return out


def optimal_fft_size(klen: int) -> int:
    """Segment FFT size for `klen` taps (include/fftconv/fftconv.hpp:53-66)."""
    return int(_capi.lib().fftconv_cuda_optimal_fft_size(int(klen)))


def add_(dst, src) -> None:
    """dst += src on the device (tail hand-over between chunks of a long signal)."""
    D, S = _Buf(dst, "dst"), _Buf(src, "src")
    if not (D.is_cuda and S.is_cuda) or D.dtype != S.dtype or D.shape != S.shape or D.ndim != 1:
        raise RuntimeError("add_: two 1-D CUDA tensors of one dtype and length expected")
    fn = getattr(_capi.lib(), f"fftconv_cuda_add_{_SFX[D.dtype]}")
    with torch.cuda.device(D.obj.device):
        _capi.check(fn(D.ptr, S.ptr, D.shape[0], _stream_of(D)))


class Filter:
    """Taps and their spectrum, resident on the current device.
    The reference transforms the kernel on every call (include/fftconv/fftconv.hpp:391-392);
    a Filter does it once.  `k` may be a numpy array or a CUDA tensor."""

    def __init__(self, k):
        K = _Buf(k, "k")
        if K.ndim != 1 or K.shape[0] == 0:
            raise RuntimeError("Number of dimensions must be one")
        self.dtype, self.klen = K.dtype, K.shape[0]
        self._sfx = _SFX[K.dtype]
        self._h = _capi.C.c_void_p()
        with _device_of(K):
            _capi.check(getattr(_capi.lib(), f"fftconv_cuda_filter_create_{self._sfx}")(
                K.ptr, self.klen, _capi.C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value and _capi is not None and _capi.lib is not None:
            _capi.lib().fftconv_cuda_filter_destroy(self._h)
            self._h = None

    __del__ = close

    def oaconvolve_(self, a, out, mode: str = "full") -> None:
        """oaconvolve_(a, k, out, mode) with this filter's taps; no spectrum kernel is launched."""
        A, O = _Buf(a, "a"), _Buf(out, "out")
        if A.ndim not in (1, 2) or O.ndim != A.ndim:
            raise RuntimeError("Number of dimensions must be one")
        if A.dtype != self.dtype or O.dtype != self.dtype:
            raise TypeError("a, out and the filter must share one dtype")
        n = A.shape[-1]
        if n < self.klen:
            raise RuntimeError("Kernel size must be smaller than input size")
        batch = 1 if A.ndim == 1 else A.shape[0]
        if A.ndim == 2 and O.shape[0] != batch:
            raise RuntimeError("out must have one row per input row")
        if A.is_cuda != O.is_cuda:
            raise RuntimeError("a and out must live on the same side (host or device)")
        fn = getattr(_capi.lib(), f"fftconv_cuda_filter_oaconvolve_{self._sfx}")
        with _device_of(A):
            _capi.check(fn(self._h, A.ptr, batch, n, A.row_stride, O.ptr, O.shape[-1], O.row_stride,
                           _mode(mode), _stream_of(A)))

    def oaconvolve(self, a, mode: str = "full"):
        n = a.shape[-1]
        out = _empty_like_rows(a, n if _mode(mode) == _capi.SAME else n + self.klen - 1)
        self.oaconvolve_(a, out, mode)
        return out

    def stream(self, rows: int = 1) -> "Stream":
        return Stream(self, rows)


class Stream:
    """Overlap-add carried ACROSS calls: `rows` signals that arrive in pieces.  push(chunk)
    returns the next chunk.shape[-1] samples of the running Full convolution of each row,
    flush() the last K-1 samples (and resets the state).  Concatenating every push and the
    flush equals oaconvolve(whole signal, k, "full").  CUDA tensors only."""

    def __init__(self, filt: Filter, rows: int = 1):
        self.filter, self.rows = filt, int(rows)
        self._h = _capi.C.c_void_p()
        _capi.check(getattr(_capi.lib(), f"fftconv_cuda_stream_create_{filt._sfx}")(
            filt._h, self.rows, _capi.C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value and _capi is not None and _capi.lib is not None:
            _capi.lib().fftconv_cuda_stream_destroy(self._h)
            self._h = None

    __del__ = close

    def _rows2d(self, x, name):
        X = _Buf(x, name)
        if not X.is_cuda or X.dtype != self.filter.dtype:
            raise RuntimeError(f"{name}: CUDA array of the filter's dtype expected")
        rows = 1 if X.ndim == 1 else X.shape[0]
        if X.ndim not in (1, 2) or rows != self.rows:
            raise RuntimeError(f"{name}: expected {self.rows} row(s)")
        return X

    def push_(self, chunk, out) -> None:
        A, O = self._rows2d(chunk, "chunk"), self._rows2d(out, "out")
        if A.shape[-1] != O.shape[-1]:
            raise RuntimeError("out must have the chunk's length")
        fn = getattr(_capi.lib(), f"fftconv_cuda_stream_push_{self.filter._sfx}")
        with _device_of(A):
            _capi.check(fn(self._h, A.ptr, A.shape[-1], A.row_stride, O.ptr, O.row_stride, _stream_of(A)))

    def push(self, chunk):
        out = _empty_like_rows(chunk, chunk.shape[-1])
        self.push_(chunk, out)
        return out

    def flush_(self, tail_out) -> None:
        O = self._rows2d(tail_out, "tail_out")
        if O.shape[-1] != self.filter.klen - 1:
            raise RuntimeError("tail_out must hold K-1 samples per row")
        fn = getattr(_capi.lib(), f"fftconv_cuda_stream_flush_{self.filter._sfx}")
        with _device_of(O):
            _capi.check(fn(self._h, O.ptr, O.row_stride, _stream_of(O)))

    def flush(self, like):
        """`like`: any CUDA tensor of the stream's device/dtype (shape donor for the result)."""
        out = _empty_like_rows(like, self.filter.klen - 1)
        self.flush_(out)
        return out


class reserved_sms:
    """Context: the persistent overlap-add kernel leaves `n` SMs free while a collective runs beside it
    (fftconv_cuda_set_reserved_sms)."""

    def __init__(self, n: int):
        self.n = int(n)

    def __enter__(self):
        self.prev = int(_capi.lib().fftconv_cuda_set_reserved_sms(self.n))

    def __exit__(self, *exc):
        _capi.lib().fftconv_cuda_set_reserved_sms(self.prev)


def oa_tail_(last, k, tail) -> None:
    """tail[i] = sum_{j > i} k[j] last[K - 1 + i - j]: the K-1 outputs past the end of a chunk's Full
    convolution, from the chunk's last K-1 samples (1-D CUDA tensors)."""
    Lb, K, Tb = _Buf(last, "last"), _Buf(k, "k"), _Buf(tail, "tail")
    klen = K.shape[0]
    if not (Lb.is_cuda and K.is_cuda and Tb.is_cuda) or Lb.shape != (klen - 1,) or Tb.shape != (klen - 1,):
        raise RuntimeError("oa_tail_: CUDA tensors of K-1, K and K-1 samples expected")
    if Lb.dtype != K.dtype or Tb.dtype != K.dtype:
        raise TypeError("last, k and tail must share one dtype")
    fn = getattr(_capi.lib(), f"fftconv_cuda_oa_tail_{_SFX[K.dtype]}")
    with _device_of(Lb):
        _capi.check(fn(Lb.ptr, K.ptr, klen, Tb.ptr, _stream_of(Lb)))


def release_stream(stream=None) -> None:
    """Free the workspace the library keeps for `stream` (a torch.cuda.Stream; default: the current one)."""
    st = torch.cuda.current_stream() if stream is None else stream
    _capi.check(_capi.lib().fftconv_cuda_release_stream(st.cuda_stream))


class MultiGpu:
    """One process driving several GPUs through the C ABI (`fftconv_cuda_mgpu_*`): batches are split
    by rows with no collective; one long signal is split into contiguous chunks whose K-1 sample
    tails go to the neighbouring GPU by ncclSend / ncclRecv.  The one-process-
    per-GPU form over `torch.distributed` is `fftconv_b200.multigpu`."""

    def __init__(self, devices=None):
        n = torch.cuda.device_count() if devices is None else len(devices)
        ids = list(range(n)) if devices is None else [int(d) for d in devices]
        arr = (_capi.C.c_int * n)(*ids)
        self._h = _capi.C.c_void_p()
        _capi.check(_capi.lib().fftconv_cuda_mgpu_init(n, arr, _capi.C.byref(self._h)))
        self.devices = ids

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value and _capi is not None and _capi.lib is not None:
            _capi.lib().fftconv_cuda_mgpu_destroy(self._h)
            self._h = None

    __del__ = close

    def sync(self) -> None:
        _capi.check(_capi.lib().fftconv_cuda_mgpu_sync(self._h))

    def chunk_bounds(self, n: int, d: int):
        lo, hi = _capi.C.c_size_t(), _capi.C.c_size_t()
        _capi.lib().fftconv_cuda_mgpu_chunk_bounds(n, len(self.devices), d, _capi.C.byref(lo), _capi.C.byref(hi))
        return int(lo.value), int(hi.value)

    def oaconvolve_batched(self, a, k, mode: str = "full"):
        """Host arrays: rows of `a` (2-D numpy) sharded over the GPUs."""
        a = np.ascontiguousarray(a)
        k = np.ascontiguousarray(k, dtype=a.dtype)
        n, klen = a.shape[1], k.shape[0]
        m = _mode(mode)
        out = np.empty((a.shape[0], n if m == _capi.SAME else n + klen - 1), dtype=a.dtype)
        fn = getattr(_capi.lib(), f"fftconv_cuda_mgpu_oaconvolve_batched_{_SFX[a.dtype.name]}")
        _capi.check(fn(self._h, a.ctypes.data, a.shape[0], n, n, k.ctypes.data, klen, out.ctypes.data,
                       out.shape[1], out.shape[1], m))
        return out

    def hilbert_batched(self, x):
        x = np.ascontiguousarray(x)
        out = np.empty_like(x)
        fn = getattr(_capi.lib(), f"fftconv_cuda_mgpu_hilbert_batched_{_SFX[x.dtype.name]}")
        _capi.check(fn(self._h, x.ctypes.data, x.shape[0], x.shape[1], x.shape[1], out.ctypes.data, x.shape[1]))
        return out

    def oaconvolve_long_host(self, a, k):
        """Host arrays: one long 1-D signal, Full mode."""
        a = np.ascontiguousarray(a)
        k = np.ascontiguousarray(k, dtype=a.dtype)
        out = np.empty(a.shape[0] + k.shape[0] - 1, dtype=a.dtype)
        fn = getattr(_capi.lib(), f"fftconv_cuda_mgpu_oaconvolve_long_host_{_SFX[a.dtype.name]}")
        _capi.check(fn(self._h, a.ctypes.data, a.shape[0], k.ctypes.data, k.shape[0], out.ctypes.data))
        return out

    def oaconvolve_long_(self, a_chunks, k, out_chunks) -> None:
        """Device-resident: `a_chunks[d]` / `out_chunks[d]` are 1-D CUDA tensors on device d; out chunk d
        holds len(a_chunks[d]) + K - 1 samples.  `k` is a host (numpy) array.  Enqueues; `sync()` waits.
        The inputs must be complete (synchronise their producers first)."""
        nd = len(self.devices)
        k = np.ascontiguousarray(k)
        sfx = _SFX[k.dtype.name]
        klen = k.shape[0]
        if len(a_chunks) != nd or len(out_chunks) != nd:
            raise RuntimeError("one chunk per device expected")
        for d in range(nd):
            if a_chunks[d].device.index != self.devices[d] or out_chunks[d].device.index != self.devices[d]:
                raise RuntimeError(f"chunk {d} must live on device {self.devices[d]}")
            if out_chunks[d].numel() < a_chunks[d].numel() + klen - 1:
                raise RuntimeError("out chunk must hold len(a chunk) + K - 1 samples")
        vp = _capi.C.c_void_p
        ap = (vp * nd)(*[int(t.data_ptr()) for t in a_chunks])
        op = (vp * nd)(*[int(t.data_ptr()) for t in out_chunks])
        ln = (_capi.C.c_size_t * nd)(*[int(t.numel()) for t in a_chunks])
        fn = getattr(_capi.lib(), f"fftconv_cuda_mgpu_oaconvolve_long_{sfx}")
        _capi.check(fn(self._h, ap, ln, k.ctypes.data, klen, op))


def launch_count() -> int:
    return int(_capi.lib().fftconv_cuda_launch_count())


def clear_cache() -> None:
    _capi.check(_capi.lib().fftconv_cuda_clear_cache())
