"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checkers for the convolution hot path of kwsp/fftconv:

* ``oracle.fftconv_oracle`` -- NumPy restatement of the reference algorithm
  (segment-size table, overlap-add, full-FFT convolution, Hilbert envelope) plus
  float64 ground truth by the mathematical definition.
* ``oracle.ref`` -- ctypes loader for ``oracle/_ref/libfftconv_ref.so``: the
  reference's own headers, unmodified, compiled over ``oracle/fftw_shim.cpp``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  The product (``fftconv_b200``) never does.
"""
