"""Counter-based synthetic signals: numpy twin of
`fftconv_b200/csrc/synthetic.cuh`, bit-identical to `fftconv_cuda_fill_synthetic_*`.

    x[i] = irwin_hall4(splitmix64(seed, i)) * scale        i = global element index

Any sample of a signal that only ever lived in HBM can be recomputed on the host, so the
measurement harness can check, exactly, outputs whose inputs sit on another GPU (the K-1
samples either side of every cut of a long signal, BASELINE.json config 5) or beyond
index 2^31.  This module is measurement / verification plumbing: nothing on the product
path imports it.
"""
from __future__ import annotations

import numpy as np

from . import _capi

SCALE = 2.6428997921303014e-05  # 1 / sqrt((65536^2 - 1) / 3): unit variance (synthetic.cuh kScale)
SEED_SIGNAL = 0x5EED0001
SEED_TAPS = 0x5EED0002

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def hash64(seed: int, idx) -> np.ndarray:
    """splitmix64 of (seed, counter); `idx` any integer array (taken modulo 2^64)."""
    i = np.asarray(idx).astype(np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (i + np.uint64(1)) * _G
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def samples(seed: int, idx, dtype=np.float32, gain: float = 1.0) -> np.ndarray:
    """x[idx] for the signal of `seed` -- the values the device generator writes."""
    h = hash64(seed, idx)
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))).astype(np.int64)
    v = (s - 131070).astype(dtype)
    return v * np.dtype(dtype).type(SCALE * gain)


def signal(seed: int, first: int, n: int, dtype=np.float32, gain: float = 1.0) -> np.ndarray:
    return samples(seed, np.arange(first, first + n, dtype=np.uint64), dtype, gain)


def fill_(x, seed: int, first: int = 0, gain: float = 1.0) -> None:
    """Fill a CUDA tensor (any shape, contiguous) with x.flat[i] = sample(seed, first + i)."""
    import torch

    assert x.is_cuda and x.is_contiguous()
    sfx = {torch.float32: "f32", torch.float64: "f64"}[x.dtype]
    fn = getattr(_capi.lib(), f"fftconv_cuda_fill_synthetic_{sfx}")
    with torch.cuda.device(x.device):
        _capi.check(fn(x.data_ptr(), x.numel(), seed, first, float(gain),
                       torch.cuda.current_stream(x.device).cuda_stream))


def conv_at(seed: int, n: int, taps, idx, *, row_first: int = 0, dtype=np.float32, gain: float = 1.0) -> np.ndarray:
    """Full-mode linear convolution of the synthetic signal x[0..n) (global indices
    row_first .. row_first + n) with `taps`, evaluated in float64 from the definition at the
    output indices `idx`:  y[o] = sum_j taps[j] x[o - j], 0 <= o - j < n."""
    taps = np.asarray(taps, dtype=np.float64)
    idx = np.asarray(idx, dtype=np.int64)
    K = taps.shape[0]
    out = np.zeros(idx.shape[0], dtype=np.float64)
    for lo in range(0, idx.shape[0], 1 << 15):          # bounded temporaries
        o = idx[lo:lo + (1 << 15)]
        src = o[:, None] - np.arange(K, dtype=np.int64)[None, :]
        ok = (src >= 0) & (src < n)
        x = samples(seed, np.where(ok, src, 0) + row_first, dtype, gain).astype(np.float64)
        out[lo:lo + o.shape[0]] = np.sum(np.where(ok, x, 0.0) * taps[None, :], axis=1)
    return out


def rel_l2(got, want) -> float:
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    d = float(np.linalg.norm(want))
    return float(np.linalg.norm(got - want)) / (d if d > 0 else 1.0)


def cut_window_indices(cuts, n_out: int, klen: int, extra_random: int = 4000, tail: int = 4000, seed: int = 0):
    """Output indices worth checking for a long signal cut at `cuts`: every
    index within K of each cut (the first K-1 outputs after a cut are the ones the neighbour's
    tail is added to), both ends of the signal, and `extra_random` hashed-random interior points."""
    parts = [np.arange(0, min(n_out, 2 * klen)), np.arange(max(0, n_out - tail), n_out)]
    for c in cuts:
        parts.append(np.arange(max(0, c - klen), min(n_out, c + klen)))
    if extra_random:
        parts.append((hash64(seed ^ 0xC0FFEE, np.arange(extra_random)) % np.uint64(n_out)).astype(np.int64))
    return np.unique(np.concatenate([p.astype(np.int64) for p in parts]))
