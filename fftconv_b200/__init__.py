"""fftconv_b200 -- B200 (sm_100a) implementation of kwsp/fftconv's convolution hot
path: `convolve`, `oaconvolve`, `hilbert` (+ in-place `_` forms and batches) behind
the reference's own API, through the C ABI of libfftconv_cuda.so."""
from .api import (Filter, MultiGpu, Stream, add_, clear_cache, convolve, convolve_, envelope, envelope_, fft, hilbert,
                  hilbert_, irfft, launch_count, oa_tail_, oaconvolve, oaconvolve_, optimal_fft_size, release_stream,
                  rf_envelope, rf_envelope_, rfft)

__version__ = "0.2.0"
__all__ = ["convolve", "convolve_", "oaconvolve", "oaconvolve_", "hilbert", "hilbert_",
           "optimal_fft_size", "add_", "launch_count", "clear_cache", "Filter", "Stream",
           "envelope", "envelope_", "rf_envelope", "rf_envelope_", "MultiGpu", "oa_tail_", "release_stream",
           "fft", "rfft", "irfft"]
