"""Small problems through every kernel, for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tests/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitize_smoke.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fftconv_b200 as fc  # noqa: E402
from oracle import fftconv_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


worst = 0.0
# N = 2048 kernel: fast + slow paths, chunked single row, misaligned rows, Same mode
for batch, n, k, mode in [(8, 12000, 245, "full"), (5, 9000, 246, "same"), (1, 60000, 245, "full"), (3, 3000, 410, "same")]:
    a = rng.standard_normal((batch, n)).astype(np.float32)
    taps = (rng.standard_normal(k) / np.sqrt(k)).astype(np.float32)
    y = fc.oaconvolve(t(a), t(taps), mode).cpu().numpy()
    worst = max(worst, O.rel_l2(y, O.conv_exact(a, taps, mode)))
big = t(rng.standard_normal((4, 9001)).astype(np.float32))
y = fc.oaconvolve(big[:, 1:], t(taps[:245].copy()), "full")   # rows 4 bytes off alignment
# fast engine: several sizes, float + double, overlap-add and single-FFT convolve
for dt in (np.float32, np.float64):
    for n, k in [(3000, 65), (5000, 165), (20000, 500), (40000, 1000)]:
        a = rng.standard_normal((3, n)).astype(dt)
        taps = (rng.standard_normal(k) / np.sqrt(k)).astype(dt)
        y = fc.oaconvolve(t(a), t(taps), "full").cpu().numpy()
        worst = max(worst, O.rel_l2(y, O.conv_exact(a, taps, "full")) * (1e-5 / 1e-12 if dt == np.float64 else 1))
    a = rng.standard_normal((3, 3000)).astype(dt)
    taps = rng.standard_normal(100).astype(dt)
    y = fc.convolve(t(a), t(taps), "same").cpu().numpy()
    # radix-4 fallback (N < 512), Hilbert fast / radix-4 / chirp-z
    y = fc.oaconvolve(t(a), t(taps[:20].copy()), "full").cpu().numpy()
    for n in (64, 1024, 8192, 1000, 333):
        x = rng.standard_normal((3, n)).astype(dt)
        e = fc.hilbert(t(x)).cpu().numpy()
        worst = max(worst, O.rel_l2(e, O.hilbert_exact(x)) * (1e-5 / 1e-12 if dt == np.float64 else 1))
# host-pointer pipeline
a = rng.standard_normal((40, 30000)).astype(np.float32)
taps = rng.standard_normal(245).astype(np.float32)
y = fc.oaconvolve(a, taps, "full")
worst = max(worst, O.rel_l2(y, O.conv_exact(a, taps, "full")))
torch.cuda.synchronize()
print("sanitize smoke done, worst normalised error", worst)
assert worst < 1e-5
