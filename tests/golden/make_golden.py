"""Generates tests/golden/reference_vectors.npz by running the REFERENCE's own code
(oracle/_ref/libfftconv_ref.so = /root/reference/include/fftconv/*.hpp, unmodified,
over oracle/fftw_shim.cpp) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The .npz (inputs and reference outputs, ~150 KB) is committed; tests never need
/root/reference at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

CASES = [
    # name, op, dtype, n, k
    ("oa_c1_f32", "oaconvolve", np.float32, 4352, 65),      # BASELINE config 1
    ("oa_f64_k95", "oaconvolve", np.float64, 1000, 95),     # test_fftconv.cpp:93-124
    ("oa_f32_k245", "oaconvolve", np.float32, 6000, 245),   # headline kernel, ragged last segment
    ("oa_f64_k245", "oaconvolve", np.float64, 3000, 245),
    ("oa_f32_k95_short", "oaconvolve", np.float32, 200, 95),  # test_fftconv.cpp:151-175
    ("cv_f64_13", "convolve", np.float64, 13, 13),          # test_fftconv.cpp:127-148 (n = k)
    ("cv_f64_22", "convolve", np.float64, 22, 22),          # even K: Same offset floor(K/2)
    ("cv_f32_94", "convolve", np.float32, 94, 94),
    ("cv_f64_1664_65", "convolve", np.float64, 1664, 65),   # test_script.cpp:29-34
]


def main():
    rng = np.random.default_rng(20240229)
    out = {}
    for name, op, dt, n, k in CASES:
        a = rng.standard_normal(n).astype(dt)
        taps = rng.standard_normal(k).astype(dt)
        out[f"{name}__a"] = a
        out[f"{name}__k"] = taps
        fn = getattr(ref, op)
        out[f"{name}__full"] = fn(a, taps, "full")
        out[f"{name}__same"] = fn(a, taps, "same")
    for name, dt, n in [("hb_f64_10", np.float64, 10), ("hb_f32_64", np.float32, 64),
                        ("hb_f64_1000", np.float64, 1000), ("hb_f32_8192", np.float32, 8192),
                        ("hb_f64_5", np.float64, 5)]:
        x = rng.standard_normal(n).astype(dt)
        out[f"{name}__x"] = x
        out[f"{name}__env"] = ref.hilbert(x)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
