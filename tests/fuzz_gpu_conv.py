"""Randomized differential run of oaconvolve / convolve on a GPU box (not collected by pytest):
random dtype, batch, n, K (around every dispatch boundary), mode, entry point, host or device pointers,
rows at every alignment with NaN guard bands, small-job rule on or off; reference = float64 direct /
FFT convolution.  `SEED=1 FUZZ_SECONDS=60 python tests/fuzz_gpu_conv.py` prints one JSON line per
failure and a summary.  Round 1: 10 524 cases, 0 failures."""
import os, sys, json, time
import numpy as np, torch
from scipy import signal
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc
dev = torch.device("cuda")
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
TOL = {np.float32: 1e-5, np.float64: 1e-12}
def rel(y, r):
    return float(np.linalg.norm(y.astype(np.float64) - r) / max(np.linalg.norm(r), 1e-300))
bad = 0; n_cases = 0; t0 = time.time()
while time.time() - t0 < float(os.environ.get("FUZZ_SECONDS", "45")):
    dt = np.float32 if rng.random() < 0.5 else np.float64
    batch = int(rng.integers(1, 10))
    n = int(rng.choice([rng.integers(1, 64), rng.integers(64, 5000), rng.integers(5000, 70000), rng.integers(70000, 700000)]))
    kmax = min(n, int(rng.choice([8, 16, 17, 64, 256, 410, 411, 1024, 1025, 3000])))
    K = int(rng.integers(1, kmax + 1))
    if batch * n * 1.0 > 3e6: batch = 1
    mode = "full" if rng.random() < 0.5 else "same"
    off_in, off_out = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    pad_in, pad_out = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    a = rng.standard_normal((batch, n)).astype(dt); k = (rng.standard_normal(K) / np.sqrt(K)).astype(dt)
    full = signal.fftconvolve(a.astype(np.float64), k.astype(np.float64)[None, :], mode="full", axes=-1) if n * K > 4e6 else np.stack([np.convolve(r.astype(np.float64), k.astype(np.float64)) for r in a])
    want = full if mode == "full" else full[:, K // 2:K // 2 + n]
    out_len = want.shape[1]
    os.environ.pop("FFTCONV_CUDA_NO_SMALL_DIRECT", None)
    if rng.random() < 0.5: os.environ["FFTCONV_CUDA_NO_SMALL_DIRECT"] = "1"
    fn = fc.oaconvolve_ if rng.random() < 0.7 else fc.convolve_
    if rng.random() < 0.6:
        tdt = torch.float32 if dt == np.float32 else torch.float64
        istore = torch.zeros(batch * (n + pad_in) + off_in, dtype=tdt, device=dev)
        iv = istore[off_in:].view(batch, n + pad_in)[:, :n]; iv.copy_(torch.from_numpy(a))
        ostore = torch.full((batch * (out_len + pad_out) + off_out,), float("nan"), dtype=tdt, device=dev)
        ov = ostore[off_out:].view(batch, out_len + pad_out)
        fn(iv, torch.from_numpy(k).to(dev), ov[:, :out_len], mode)
        torch.cuda.synchronize()
        y = ov[:, :out_len].cpu().numpy()
        guard_ok = bool(torch.isnan(ov[:, out_len:]).all()) and bool(torch.isnan(ostore[:off_out]).all())
    else:
        y = np.full((batch, out_len), np.nan, dt); fn(a, k, y, mode); guard_ok = True
    e = rel(y, want); n_cases += 1
    if not (e < TOL[dt] and np.isfinite(y).all() and guard_ok):
        bad += 1
        print(json.dumps({"FAIL": True, "dtype": dt.__name__, "batch": batch, "n": n, "K": K, "mode": mode, "fn": fn.__name__, "err": e, "guard": guard_ok, "off": [off_in, off_out, pad_in, pad_out], "small_direct_off": os.environ.get("FFTCONV_CUDA_NO_SMALL_DIRECT")}), flush=True)
print(json.dumps({"cases": n_cases, "failures": bad}))
