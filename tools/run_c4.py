"""Time BASELINE config 4 (Hilbert envelope, 8192 x 8192 float32) on the current device: the real-input
engine against the register-blocked one.  usage: python tools/run_c4.py [steps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
x = torch.randn((8192, 8192), device=dev)
out = torch.empty_like(x)


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        fc.hilbert_(x, out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fc.hilbert_(x, out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
variants = [("r2c", "0", None), ("register_blocked", "1", None)]
if len(sys.argv) > 2:
    variants += [(f"r2c_groups{g}", "0", g) for g in sys.argv[2].split(",")]
for name, env, groups in variants:
    os.environ["FFTCONV_CUDA_NO_HR2C"] = env
    os.environ.pop("FFTCONV_CUDA_HR2C_GROUPS", None)
    if groups:
        os.environ["FFTCONV_CUDA_HR2C_GROUPS"] = groups
    ms = timed(steps)
    ms_long = timed(3000)
    ref = np.abs(__import__("scipy.signal").signal.hilbert(x[:4].cpu().numpy().astype(np.float64), axis=-1))
    fc.hilbert_(x, out)
    err = float(np.linalg.norm(out[:4].cpu().numpy() - ref) / np.linalg.norm(ref))
    res[name] = {"ms": ms, "ms_3000": ms_long, "frac_of_6547": 536870912 / (ms * 1e-3) / 1e9 / 6547.5, "rel_l2": err}
print(json.dumps(res))
