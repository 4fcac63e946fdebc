"""A/B timing of BASELINE config 2 (4096 x 65536 float32, K taps) under environment switches of the library.
usage: python tools/ab_c2.py "NAME=ENV1=v,ENV2=v" ... [--k 245] [--rows 4096] [--long 2000]
Each variant: 200 back-to-back steps and `long` steps (sustained, power-capped) between CUDA events."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftconv_b200 as fc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("variants", nargs="+")
ap.add_argument("--k", type=int, default=245)
ap.add_argument("--rows", type=int, default=4096)
ap.add_argument("--long", type=int, default=2000)
ap.add_argument("--flush", action="store_true", help="also: 200 steps timed one by one with an L2 flush (256 MB write) before each")
args = ap.parse_args()
dev = torch.device("cuda:0")
a = torch.randn((args.rows, 65536), device=dev)
k = torch.randn(args.k, device=dev) / args.k ** 0.5
out = torch.empty((args.rows, 65536 + args.k - 1), device=dev)


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        fc.oaconvolve_(a, k, out, "full")
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fc.oaconvolve_(a, k, out, "full")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush else None


def timed_flushed(n):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        flush_buf.zero_()
        e0.record()
        fc.oaconvolve_(a, k, out, "full")
        e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in ev) / n


ref = None
for spec in args.variants:
    name, _, envs = spec.partition("=")
    sets = dict(e.split("=") for e in envs.split(",") if e)
    for kk, vv in sets.items():
        os.environ[kk] = vv
    ms = timed(200)
    ms_long = timed(args.long)
    ms_flushed = timed_flushed(200) if args.flush else None
    fc.oaconvolve_(a, k, out, "full")
    torch.cuda.synchronize()
    chk = out[:2].double().cpu()
    if ref is None:
        ref = chk
    err = float((chk - ref).norm() / ref.norm())
    print(json.dumps({"variant": name, "env": sets, "k": args.k, "rows": args.rows, "ms_200": ms, "ms_long": ms_long, "ms_flushed": ms_flushed,
                      "rel_l2_vs_first_variant": err}), flush=True)
    for kk in sets:
        os.environ.pop(kk, None)
