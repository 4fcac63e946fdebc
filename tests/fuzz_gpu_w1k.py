"""Randomized differential run of the warp-per-FFT overlap-add kernel (oa_w1k_*.cuh) on a GPU box (not collected by
pytest): the engine is forced (FFTCONV_CUDA_OA_FFT_SIZE=1024), random batch, row length (short rows, one block,
many blocks), K in 8 .. 400, mode, row padding and base misalignment with NaN guard bands, grid cut down to a random
number of SMs (long ranges per warp), dynamic tail on or off, filter handle or plain call; reference = float64
convolution.  `SEED=1 FUZZ_SECONDS=60 python tests/fuzz_gpu_w1k.py` prints one JSON line per failure and a summary."""
import json
import os
import sys
import time

import numpy as np
import torch
from scipy import signal

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc  # noqa: E402
from fftconv_b200 import api  # noqa: E402

dev = torch.device("cuda")
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
os.environ["FFTCONV_CUDA_OA_FFT_SIZE"] = "1024"
sms = torch.cuda.get_device_properties(0).multi_processor_count
bad = 0
cases = 0
t0 = time.time()
while time.time() - t0 < float(os.environ.get("FUZZ_SECONDS", "45")):
    K = int(rng.choice([8, 9, 17, 33, 64, 65, 129, 245, 246, 300, 360, 400, int(rng.integers(8, 401))]))
    n = int(rng.choice([rng.integers(K, 1100), rng.integers(1100, 9000), rng.integers(9000, 140000)]))
    batch = int(rng.integers(1, max(2, min(600, 3_000_000 // n))))
    mode = "full" if rng.random() < 0.5 else "same"
    off_in, off_out = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    pad_in, pad_out = int(rng.integers(0, 5)), int(rng.integers(0, 5))
    a = rng.standard_normal((batch, n)).astype(np.float32)
    k = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    full = signal.fftconvolve(a.astype(np.float64), k.astype(np.float64)[None, :], mode="full", axes=-1)
    want = full if mode == "full" else full[:, K // 2:K // 2 + n]
    out_len = want.shape[1]
    istore = torch.zeros(batch * (n + pad_in) + off_in, dtype=torch.float32, device=dev)
    iv = istore[off_in:].view(batch, n + pad_in)[:, :n]
    iv.copy_(torch.from_numpy(a))
    ostore = torch.full((batch * (out_len + pad_out) + off_out + 8,), float("nan"), dtype=torch.float32, device=dev)
    ov = ostore[off_out:off_out + batch * (out_len + pad_out)].view(batch, out_len + pad_out)
    pool = int(rng.choice([0, 0, 20, 60]))
    os.environ["FFTCONV_CUDA_W1K_POOL_PCT"] = str(pool)
    keep = int(rng.choice([0, 0, sms - 1, sms - 3, sms - 20]))
    use_filter = rng.random() < 0.3
    with api.reserved_sms(keep):
        reps = 2 if pool else 1
        for _ in range(reps):
            ostore.fill_(float("nan"))
            if use_filter:
                f = fc.Filter(k)
                f.oaconvolve_(iv, ov[:, :out_len], mode) if hasattr(f, "oaconvolve_") else ov[:, :out_len].copy_(f.oaconvolve(iv, mode))
                f.close()
            else:
                fc.oaconvolve_(iv, torch.from_numpy(k).to(dev), ov[:, :out_len], mode)
            torch.cuda.synchronize()
            got = ov[:, :out_len].cpu().numpy().astype(np.float64)
            err = float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300))
            guards_ok = bool(torch.isnan(ov[:, out_len:]).all()) and bool(torch.isnan(ostore[:off_out]).all()) and \
                bool(torch.isnan(ostore[off_out + batch * (out_len + pad_out):]).all())
            cases += 1
            if not (err < 1e-5) or not guards_ok or np.isnan(got).any():
                bad += 1
                print(json.dumps({"K": K, "n": n, "batch": batch, "mode": mode, "off": [off_in, off_out], "pad": [pad_in, pad_out],
                                  "pool": pool, "reserved": keep, "filter": use_filter, "err": err, "guards_ok": guards_ok}), flush=True)
print(json.dumps({"cases": cases, "failures": bad}))
sys.exit(1 if bad else 0)
