"""Shared test data: the reference's literal known-answer vectors and the golden
fixtures generated from the reference's own code (tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")

# /root/reference/test/test_hilbert.cpp:8-15
HILBERT_IN = np.array([-0.999984, -0.736924, 0.511211, -0.0826997, 0.0655345,
                       -0.562082, -0.905911, 0.357729, 0.358593, 0.869386])
HILBERT_EXPECT = np.array([1.45197493, 1.15365169, 0.54703078, 0.27346519, 0.15097965,
                           0.83696245, 1.1476185, 0.71885109, 0.46089151, 1.07384968])

# /root/reference/test/test_fftw.cpp:266-285 and :352-369 (20-point DFT known answers)
DFT20_IN = np.array([0.24800501, 0.19901191, 0.22109176, 0.65214355, 0.13190115,
                     0.83696173, 0.36607799, 0.59169134, 0.89522796, 0.03245338,
                     0.39925889, 0.93391339, 0.50966463, 0.26965452, 0.61894752,
                     0.29961656, 0.4266788, 0.03522679, 0.0617605, 0.57214474])
DFT20_RE = np.array([8.30143212, -1.4786986, -0.28946925, 0.59119628, 0.17941478, 0.54434089,
                     -0.66171488, 0.48975843, 0.12947463, -0.9028664, -0.5442037])
DFT20_IM = np.array([0., -0.70254131, -0.35119795, 0.43823534, -0.23602404, 1.67620125,
                     -0.42226181, 2.11848431, -0.83078287, -1.10366614, 0.])

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}  # BASELINE.json north_star


def golden():
    return np.load(GOLDEN)


def golden_conv_cases():
    g = golden()
    names = sorted({k.split("__")[0] for k in g.files if k.startswith(("oa_", "cv_"))})
    return g, names


def golden_hilbert_cases():
    g = golden()
    names = sorted({k.split("__")[0] for k in g.files if k.startswith("hb_")})
    return g, names
