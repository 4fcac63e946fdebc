"""The warp-per-FFT overlap-add kernel (oa_w1k_*.cuh) under compute-sanitizer: fixed ranges, the dynamic tail,
filter handle, Same mode, clipped and misaligned rows -- on a grid cut down to a few SMs so that the tools finish.

    compute-sanitizer --tool memcheck  python tests/sanitize_w1k.py
    compute-sanitizer --tool racecheck python tests/sanitize_w1k.py small
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fftconv_b200 as fc  # noqa: E402
from fftconv_b200 import api  # noqa: E402
from oracle import fftconv_oracle as O  # noqa: E402

small = "small" in sys.argv
dev = torch.device("cuda:0")
rng = np.random.default_rng(5)
os.environ["FFTCONV_CUDA_OA_FFT_SIZE"] = "1024"          # the engine under test, whatever the job size
sms = torch.cuda.get_device_properties(0).multi_processor_count
keep = api.reserved_sms(sms - (2 if small else 6))        # 2 or 6 CTAs of 16 warps
keep.__enter__()


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


worst = 0.0
cases = [(13, 65536, 245, "full"), (5, 4101, 65, "same")] if small else \
    [(40, 65536, 245, "full"), (33, 8192, 245, "same"), (9, 4101, 300, "full"), (64, 2048, 17, "same")]
for pool in ("0", "40"):
    os.environ["FFTCONV_CUDA_W1K_POOL_PCT"] = pool
    for batch, n, k, mode in cases:
        a = rng.standard_normal((batch, n)).astype(np.float32)
        taps = (rng.standard_normal(k) / np.sqrt(k)).astype(np.float32)
        for _ in range(2):                                  # twice: the dynamic tail's counters must be back at zero
            y = fc.oaconvolve(t(a), t(taps), mode).cpu().numpy()
            worst = max(worst, O.rel_l2(y, O.conv_exact(a, taps, mode)))
        f = fc.Filter(taps)
        y = f.oaconvolve(t(a), mode).cpu().numpy()
        f.close()
        worst = max(worst, O.rel_l2(y, O.conv_exact(a, taps, mode)))
    off = t(rng.standard_normal((4, 9001)).astype(np.float32))
    y = fc.oaconvolve(off[:, 1:], t(taps), "full")          # rows 4 bytes off 16-byte alignment
torch.cuda.synchronize()
keep.__exit__(None, None, None)
print("sanitize w1k done, worst error", worst)
assert worst < 1e-5
