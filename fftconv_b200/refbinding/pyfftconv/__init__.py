"""The reference's own pybind11 module (/root/reference/py/main.cpp, compiled UNCHANGED
against include/fftconv/*.hpp of this repo and linked to libfftconv_cuda.so) -- the
drop-in proof for the Python boundary.  Built by fftconv_b200.build.build_ref_binding()
where the reference tree exists; the .so is git-ignored and ships with the snapshot."""
from ._pyfftconv import (__doc__, __version__, convolve, convolve_, hilbert, hilbert_,  # noqa: F401
                         oaconvolve, oaconvolve_)
