"""Time overlap-add on a few shapes of equal total size (float32, K = 245): long rows against short rows,
Full against Same.  usage: python tools/time_oa_shapes.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc  # noqa: E402

dev = torch.device("cuda:0")
dtype = torch.float64 if "f64" in sys.argv else torch.float32
K = int(sys.argv[sys.argv.index("--k") + 1]) if "--k" in sys.argv else 245
k = torch.randn(K, device=dev, dtype=dtype) / K ** 0.5
for rows, n in ((1024, 65536), (8192, 8192), (16384, 4096), (32768, 2048)):
    if dtype == torch.float64:
        rows //= 2
    a = torch.randn((rows, n), device=dev, dtype=dtype)
    for mode in ("full", "same"):
        out = torch.empty((rows, n + K - 1 if mode == "full" else n), device=dev, dtype=dtype)
        for _ in range(5):
            fc.oaconvolve_(a, k, out, mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            fc.oaconvolve_(a, k, out, mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        print(json.dumps({"rows": rows, "n": n, "mode": mode, "ms": ms,
                          "dtype": str(dtype), "k": K,
                          "frac_of_6547": rows * (n + out.shape[1]) * a.element_size() / (ms * 1e-3) / 1e9 / 6547.5}), flush=True)
