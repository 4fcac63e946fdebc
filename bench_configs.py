#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py).

    python bench_configs.py                     # c1, c3, c4, c5 on one GPU -> one JSON line each
    torchrun --nproc-per-node N bench_configs.py --config c5 --gpus N    # long signal across N GPUs

Each line: device-resident throughput (CUDA events, >= 3 warm-ups, inputs larger than L2
or L2 flushed between iterations), algorithmic bytes per launch, and the
fraction of the measured HBM copy peak.  A bounded CPU reference leg (oracle/_ref) is timed for
c1/c3/c4 on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_cuda(fn, steps, warmup, flush=None):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush is None:
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    for _ in range(steps):
        flush()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / steps


def cpu_leg(kind, dt, rows, n, k, seconds=8.0):
    import numpy as np

    from oracle import ref

    if not ref.available():
        return None
    rng = np.random.default_rng(1)
    threads = ref.hardware_threads()
    a = rng.standard_normal((max(threads, 8), n)).astype(dt)
    taps = rng.standard_normal(k).astype(dt) if k else None
    fn = {"oa": lambda x: ref.oaconvolve(x, taps, "full"), "conv": lambda x: ref.convolve(x, taps, "full"),
          "hilbert": lambda x: ref.hilbert(x)}[kind]
    fn(a)
    t0 = time.perf_counter()
    out = fn(a)
    tp = time.perf_counter() - t0
    reps = int(max(1, min(rows / a.shape[0], seconds / max(tp, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn(a)
    dt_s = (time.perf_counter() - t0) / reps
    return {"value": out.size / dt_s / 1e9, "unit": "Gsamples/s", "cores": threads, "kind": "reference",
            "sample": f"{a.shape[0]} rows x {n}, {reps} repetitions; reference headers over oracle/fftw_shim.cpp"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2n", type=int, default=31, help="c5 signal length = 2**log2n")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    import fftconv_b200 as fc
    from fftconv_b200 import multigpu as mg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = peak_gbs()
    which = ["c1", "c3", "c4", "c5"] if args.config == "all" else [args.config]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush():
        flush_buf.zero_()

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    if "c1" in which and world == 1:
        # oaconvolve float32, n = 4352, k = 65 (test_script.cpp:29-34; README.md:99-102: 17.2 us/call on an i7)
        a = torch.randn(4352, device=dev, generator=g)
        k = torch.randn(65, device=dev, generator=g)
        out = torch.empty(4352 + 64, device=dev)
        ms = time_cuda(lambda: fc.oaconvolve_(a, k, out, "full"), 200, 20)
        # the same call replayed from a CUDA graph: device time per call without the Python / launch overhead
        graph_us = None
        try:
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                fc.oaconvolve_(a, k, out, "full")
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    for _ in range(50):
                        fc.oaconvolve_(a, k, out, "full")
            torch.cuda.synchronize()
            graph_us = time_cuda(gr.replay, 20, 5) * 1e3 / 50
        except Exception as e:  # capture not possible: report only the eager number
            graph_us = None
            torch.cuda.synchronize()
        an, kn = a.cpu().numpy(), k.cpu().numpy()
        on = np.empty(4416, np.float32)
        for _ in range(20):
            fc.oaconvolve_(an, kn, on, "full")
        t0 = time.perf_counter()
        for _ in range(200):
            fc.oaconvolve_(an, kn, on, "full")
        host_us = (time.perf_counter() - t0) / 200 * 1e6
        emit({"config": "c1", "workload": "oaconvolve f32 n=4352 k=65 full, single signal (launch-bound)",
              "device_us_per_call": ms * 1e3, "device_us_per_call_cuda_graph": graph_us,
              "host_api_us_per_call": host_us, "value": 4416 / (ms * 1e-3) / 1e9,
              "unit": "Gsamples/s", "algorithmic_bytes": 35332,
              "published_cpu_us_per_call": 17.2, "published_source": "README.md:99-102 (i7, FFTW, 1 thread)",
              "cpu_baseline": None if args.no_cpu else cpu_leg("oa", np.float32, 4096, 4352, 65, 4.0)})

    if "c3" in which and world == 1:
        a = torch.randn((1024, 16384), device=dev, generator=g, dtype=torch.float64)
        k = torch.randn(165, device=dev, generator=g, dtype=torch.float64)
        out = torch.empty((1024, 16384 + 164), device=dev, dtype=torch.float64)
        os.environ.pop("FFTCONV_CUDA_CONV_SINGLE_FFT", None)
        ms = time_cuda(lambda: fc.convolve_(a, k, out, "full"), args.steps, args.warmup, flush)
        os.environ["FFTCONV_CUDA_CONV_SINGLE_FFT"] = "1"
        ms_one = time_cuda(lambda: fc.convolve_(a, k, out, "full"), args.steps, args.warmup, flush)
        os.environ.pop("FFTCONV_CUDA_CONV_SINGLE_FFT", None)
        alg = 1024 * (16384 + 16548) * 8 + 165 * 8
        emit({"config": "c3", "workload": "batched convolve f64 1024 x 16384, k=165, full",
              "ms_per_step": ms, "value": out.numel() / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
              "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None},
              "algorithmic_bytes": alg, "l2": "flushed between iterations (256 MB write)",
              "method": "on-chip segment FFTs of 2048 points (default beyond one CTA's FFT; same linear convolution)",
              "alt_one_long_fft": {"ms_per_step": ms_one, "value": out.numel() / (ms_one * 1e-3) / 1e9,
                                   "frac": alg / (ms_one * 1e-3) / 1e9 / peak,
                                   "note": "FFTCONV_CUDA_CONV_SINGLE_FFT=1: one FFT of length 32768 per row pair "
                                           "(16 x 2048 through global scratch), the reference convolve_fftw's method"},
              "cpu_baseline": None if args.no_cpu else cpu_leg("conv", np.float64, 1024, 16384, 165)})

    if "c4" in which and world == 1:
        x = torch.randn((8192, 8192), device=dev, generator=g)
        env = torch.empty_like(x)
        ms = time_cuda(lambda: fc.hilbert_(x, env), args.steps, args.warmup)
        alg = 2 * 8192 * 8192 * 4
        emit({"config": "c4", "workload": "batched Hilbert envelope f32 8192 x 8192",
              "ms_per_step": ms, "value": env.numel() / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
              "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None},
              "algorithmic_bytes": alg, "l2": "inputs (268 MB) exceed L2",
              "cpu_baseline": None if args.no_cpu else cpu_leg("hilbert", np.float32, 8192, 8192, 0)})

    if "rf" in which and world == 1:
        # the chain configs 2 + 4 imply: band-pass FIR, envelope, log
        # compression, decimation -- one library call (two kernels, post-processing in the store of
        # the second) against the same steps as separate passes over HBM
        rows, n, klen, dec = 8192, 8192, 245, 4
        x = torch.randn((rows, n), device=dev, generator=g)
        k = torch.randn(klen, device=dev, generator=g) / klen ** 0.5
        out = torch.empty((rows, n // dec), device=dev)
        kw = dict(log_compress=True, ref=1.0, dynamic_range_db=60.0, decimate=dec)
        ms_chain = time_cuda(lambda: fc.rf_envelope_(x, k, out, **kw), args.steps, args.warmup)
        ms_env = time_cuda(lambda: fc.envelope_(x, out, **kw), args.steps, args.warmup)
        filt, env = torch.empty_like(x), torch.empty_like(x)

        def separate():
            fc.oaconvolve_(x, k, filt, "same")
            fc.hilbert_(filt, env)
            return torch.clamp(1.0 + 20.0 * torch.log10(env[:, ::dec]) / 60.0, 0.0, 1.0)

        ms_sep = time_cuda(separate, args.steps, args.warmup)
        ms_hil = time_cuda(lambda: fc.hilbert_(x, env), args.steps, args.warmup)
        ms_post = time_cuda(lambda: torch.clamp(1.0 + 20.0 * torch.log10(env[:, ::dec]) / 60.0, 0.0, 1.0),
                            args.steps, args.warmup)
        alg = rows * (n + n // dec) * 4 + klen * 4
        emit({"config": "rf", "workload": f"RF chain f32 {rows} x {n}: band-pass k={klen} (same) -> envelope -> "
                                          f"log compression -> decimate by {dec}",
              "ms_per_step": ms_chain, "value": rows * n / (ms_chain * 1e-3) / 1e9, "unit": "Ginput-samples/s",
              "roofline": {"bound": "hbm", "achieved": alg / (ms_chain * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": alg / (ms_chain * 1e-3) / 1e9 / peak, "traffic": None},
              "algorithmic_bytes": alg,
              "separate_passes_ms": ms_sep, "envelope_fused_ms": ms_env, "hilbert_plain_ms": ms_hil,
              "torch_postprocessing_ms": ms_post,
              "note": "separate = oaconvolve_ + hilbert_ + torch log10/clamp on a strided view"})

    if "pcie" in which and world == 1:
        # what the host-pointer path can reach at best: pinned copies alone and in both directions at once
        nbytes = 1 << 30
        hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        din = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dout = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def wall(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps

        def h2d():
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)

        def both():
            h2d()
            d2h()

        t_h2d, t_d2h, t_both = wall(h2d), wall(d2h), wall(both)
        emit({"config": "pcie", "bytes_each_way": nbytes, "h2d_gbs": nbytes / t_h2d / 1e9, "d2h_gbs": nbytes / t_d2h / 1e9,
              "both_gbs_each_way": nbytes / t_both / 1e9,
              "c2_e2e_floor_ms": max(1073742804, 1077739520) / (nbytes / t_both) * 1e3})

    if "ksweep" in which and world == 1:
        # batched overlap-add at the headline shape for one K per row of the segment-size table
        # (include/fftconv/fftconv.hpp:38-49): which kernel serves it and how far from the roofline
        rows, n = 4096, 65536
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            r = rows if dt == torch.float32 else rows // 2
            a = torch.randn((r, n), device=dev, generator=g, dtype=dt)
            for klen in (6, 20, 35, 65, 165, 245, 500, 768, 1000, 3000):
                k = torch.randn(klen, device=dev, generator=g, dtype=dt)
                out = torch.empty((r, n + klen - 1), device=dev, dtype=dt)
                try:
                    ms = time_cuda(lambda: fc.oaconvolve_(a, k, out, "full"), max(3, args.steps // 2), 3)
                except RuntimeError as e:
                    emit({"config": "ksweep", "dtype": name, "k": klen, "error": str(e)[:120]})
                    continue
                alg = (a.numel() + out.numel() + klen) * a.element_size()
                emit({"config": "ksweep", "dtype": name, "k": klen, "fft_size": fc.optimal_fft_size(klen), "rows": r,
                      "n": n, "ms_per_step": ms, "value": out.numel() / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                      "frac_of_measured_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak})
            del a

    if "c5" in which:
        n = 1 << args.log2n
        klen = 245
        lo, hi = mg.chunk_bounds(n, world, rank)
        a = torch.randn(hi - lo, device=dev, generator=g)
        k = torch.randn(klen, device=torch.device("cpu"), generator=torch.Generator().manual_seed(7)).to(dev)

        def step():
            return mg.oaconvolve_long(a, k, rank=rank, world=world)

        for _ in range(max(args.warmup, 3)):
            y = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            y = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # sampled exact check on this rank's slice (float64 dot products)
        from oracle import fftconv_oracle as O

        m = hi - lo
        win = min(m, 1 << 22)                     # exact check inside the first 4 Mi samples of the chunk
        idx = np.unique(np.concatenate([np.arange(klen - 1, min(win, 4000)),
                                        np.random.default_rng(rank).integers(klen - 1, win, 4000),
                                        np.arange(max(klen - 1, win - 4000), win)]))
        want = O.conv_exact_at(a[:win].cpu().numpy(), k.cpu().numpy(), idx)
        got = y[:win].cpu().numpy()[idx]
        err = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        alg = (n + n + klen - 1) * 4 + klen * 4
        emit({"config": "c5", "workload": f"single long signal f32 n=2^{args.log2n}, k=245, full, {world} GPU(s)",
              "n_gpus": world, "ms_per_step": ms, "value": (n + klen - 1) / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
              "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                           "frac": alg / (ms * 1e-3) / 1e9 / (peak * world), "traffic": None},
              "algorithmic_bytes": alg, "tail_exchange_bytes_per_cut": (klen - 1) * 4,
              "interior_rel_l2_sampled": err})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
