"""Randomized differential run of hilbert / envelope on a GPU box (not collected by pytest): any
length up to 131072 (float32) / 65536 (float64), powers of two and not, decimation and log
compression, host or device pointers; reference = scipy.signal.hilbert in float64.  Round 2: one case in
seven is a batch of 64 .. 700 float32 lines of 8192 samples (the real-input engine, hilbert_r2c_*.cuh) and one
in twenty a line of up to 2^20 samples (the general FFT path, fft_global.cuh).
`SEED=3 FUZZ_SECONDS=50 python tests/fuzz_gpu_hilbert.py`.  Round 1: 3 478 cases, 0 failures."""
import os, sys, json, time
import numpy as np, torch
from scipy import signal
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc
dev = torch.device("cuda")
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
TOL = {np.float32: 1e-5, np.float64: 1e-12}
bad = 0; n_cases = 0; t0 = time.time()
while time.time() - t0 < float(os.environ.get("FUZZ_SECONDS", "45")):
    dt = np.float32 if rng.random() < 0.5 else np.float64
    nmax = 131072 if dt == np.float32 else 65536
    n = int(rng.choice([rng.integers(1, 64), rng.integers(64, 9000), 2 ** int(rng.integers(1, 17)), rng.integers(9000, nmax)]))
    batch = int(rng.integers(1, 8)) if n < 20000 else int(rng.integers(1, 3))
    u = rng.random()
    if u < 0.15:
        dt, n, batch = np.float32, 8192, int(rng.integers(64, 700))
    elif u < 0.20:
        n, batch = int(rng.integers(nmax, 1 << 20)), 1
    dec = int(rng.choice([1, 1, 2, 3, 5, 16])); log = bool(rng.random() < 0.5)
    x = (rng.standard_normal((batch, n)) * np.exp(rng.uniform(-3, 1, (batch, 1)))).astype(dt)
    env = np.abs(signal.hilbert(x.astype(np.float64), axis=-1))[:, ::dec]
    want = np.clip(1.0 + 20.0 * np.log10(np.maximum(env, 1e-300) / 1.3) / 55.0, 0.0, 1.0) if log else env
    kw = dict(log_compress=log, ref=1.3, dynamic_range_db=55.0, decimate=dec)
    try:
        if rng.random() < 0.6:
            y = fc.envelope(torch.from_numpy(x).to(dev), **kw); torch.cuda.synchronize(); y = y.cpu().numpy()
        else:
            y = fc.envelope(x, **kw)
    except RuntimeError as e:
        print(json.dumps({"ERR": str(e)[:100], "n": n, "dtype": dt.__name__})); bad += 1; continue
    n_cases += 1
    if log:
        e = float(np.linalg.norm(y.astype(np.float64) - want) / np.sqrt(want.size)); ok = e < 4 * TOL[dt]
    else:
        e = float(np.linalg.norm(y.astype(np.float64) - want) / max(np.linalg.norm(want), 1e-300)); ok = e < TOL[dt]
    if not (ok and np.isfinite(y).all() and y.shape == want.shape):
        bad += 1
        print(json.dumps({"FAIL": True, "dtype": dt.__name__, "batch": batch, "n": n, "dec": dec, "log": log, "err": e}), flush=True)
print(json.dumps({"cases": n_cases, "failures": bad}))
