"""A few launches of the config-4 Hilbert kernel for ncu.  usage: python tools/run_c4_once.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = torch.randn((8192, 8192), device="cuda:0")
out = torch.empty_like(x)
for _ in range(n):
    fc.hilbert_(x, out)
torch.cuda.synchronize()
print("ok")
