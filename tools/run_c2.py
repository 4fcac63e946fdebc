#!/usr/bin/env python
"""One short run of the headline shape (BASELINE config 2) for profiling under ncu:
    python tools/run_c2.py [steps] [rows]
Environment switches of the library (FFTCONV_CUDA_*) select the engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import fftconv_b200 as fc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
a = torch.randn((rows, 65536), device=dev)
k = torch.randn(245, device=dev) / 245 ** 0.5
out = torch.empty((rows, 65536 + 244), device=dev)
for _ in range(steps):
    fc.oaconvolve_(a, k, out, "full")
torch.cuda.synchronize()
print("ok", float(out[0, :10].abs().sum()))
