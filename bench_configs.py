#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py's headline).

    python bench_configs.py                     # c1, c3, c4, c5 on one GPU -> one JSON line each
    torchrun --nproc-per-node N bench_configs.py --config c5 --gpus N    # long signal across N GPUs

Each config is a function that returns one dict: device-resident throughput (CUDA events, >= 3
warm-ups, inputs larger than L2 or L2 flushed between iterations), algorithmic bytes per launch
(SURVEY.md section 8d), the fraction of the measured HBM copy peak, and a parity figure against
the float64 definition on the same inputs.  `bench.py` imports them for the `secondary` block of
its JSON line, so the driver's record carries them too.
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


def roofline(alg_bytes, ms, peak, world=1):
    ach = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s", "frac": ach / (peak * world),
            "traffic": None}


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


class Ctx:
    def __init__(self, dev, rank=0, world=1):
        import torch

        self.dev, self.rank, self.world = dev, rank, world
        self.peak = peak_gbs()
        self.g = torch.Generator(device=dev).manual_seed(1234 + rank)
        self._flush = None

    def flush(self):
        import torch

        if self._flush is None:
            self._flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self._flush.zero_()


def config_c1(cx, steps=200, cpu=False):
    """oaconvolve float32, n = 4352, k = 65 (test_script.cpp:29-34; README.md:99-102: 17.2 us/call on an i7)."""
    import numpy as np
    import torch

    import fftconv_b200 as fc
    from fftconv_b200 import synthetic as syn

    dev = cx.dev
    a = torch.randn(4352, device=dev, generator=cx.g)
    k = torch.randn(65, device=dev, generator=cx.g)
    out = torch.empty(4352 + 64, device=dev)
    ms = time_cuda(lambda: fc.oaconvolve_(a, k, out, "full"), steps, 20)
    graph_us = None
    try:  # the same call replayed from a CUDA graph: device time per call without Python / launch overhead
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fc.oaconvolve_(a, k, out, "full")
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                for _ in range(50):
                    fc.oaconvolve_(a, k, out, "full")
        torch.cuda.synchronize()
        graph_us = time_cuda(gr.replay, 20, 5) * 1e3 / 50
    except Exception:
        graph_us = None
        torch.cuda.synchronize()
    an, kn = a.cpu().numpy(), k.cpu().numpy()
    on = np.empty(4416, np.float32)
    for _ in range(50):
        fc.oaconvolve_(an, kn, on, "full")
    t0 = time.perf_counter()
    for _ in range(steps):
        fc.oaconvolve_(an, kn, on, "full")
    host_us = (time.perf_counter() - t0) / steps * 1e6
    want = np.convolve(an.astype(np.float64), kn.astype(np.float64))
    return {"config": "c1", "workload": "oaconvolve f32 n=4352 k=65 full, single signal (launch-bound)",
            "device_us_per_call": ms * 1e3, "device_us_per_call_cuda_graph": graph_us,
            "host_api_us_per_call_python": host_us, "host_api_us_per_call_cpp": cpp_c1_latency(),
            "host_api_path": "pageable numpy arrays -> mapped page-locked buffer -> one zero-copy kernel",
            "value": 4416 / (ms * 1e-3) / 1e9, "unit": "Gsamples/s", "algorithmic_bytes": 35332,
            "rel_l2_vs_float64_definition": {"device": syn.rel_l2(out.cpu().numpy(), want), "host_api": syn.rel_l2(on, want)},
            "published_cpu_us_per_call": 17.2, "published_source": "README.md:99-102 (i7, FFTW, 1 thread)",
            "cpu_baseline": cpu_leg("oa", np.float32, 4096, 4352, 65, 4.0) if cpu else None}


def cpp_c1_latency():
    """The reference's own benchmark case timed through the C++ span API (tests/cpp/bench_c1)."""
    import subprocess

    exe = os.path.join(ROOT, "tests", "cpp", "bench_c1")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception:
        return None


def config_c3(cx, steps=10, warmup=3, cpu=False):
    import numpy as np
    import torch

    import fftconv_b200 as fc
    from fftconv_b200 import synthetic as syn

    dev = cx.dev
    a = torch.randn((1024, 16384), device=dev, generator=cx.g, dtype=torch.float64)
    k = torch.randn(165, device=dev, generator=cx.g, dtype=torch.float64)
    out = torch.empty((1024, 16384 + 164), device=dev, dtype=torch.float64)
    os.environ.pop("FFTCONV_CUDA_CONV_SINGLE_FFT", None)
    ms = time_cuda(lambda: fc.convolve_(a, k, out, "full"), steps, warmup, cx.flush)
    rows = [0, 511, 1023]
    want = np.stack([np.convolve(a[r].cpu().numpy(), k.cpu().numpy()) for r in rows])
    err = syn.rel_l2(out[rows].cpu().numpy(), want)
    os.environ["FFTCONV_CUDA_CONV_SINGLE_FFT"] = "1"
    ms_one = time_cuda(lambda: fc.convolve_(a, k, out, "full"), steps, warmup, cx.flush)
    err_one = syn.rel_l2(out[rows].cpu().numpy(), want)
    os.environ.pop("FFTCONV_CUDA_CONV_SINGLE_FFT", None)
    alg = 1024 * (16384 + 16548) * 8 + 165 * 8
    return {"config": "c3", "workload": "batched convolve f64 1024 x 16384, k=165, full",
            "ms_per_step": ms, "value": out.numel() / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "roofline": roofline(alg, ms, cx.peak), "algorithmic_bytes": alg,
            "l2": "flushed between iterations (256 MB write)",
            "rel_l2_vs_float64_definition": err, "tolerance": 1e-12,
            "method": "on-chip segment FFTs of 2048 points (default beyond one CTA's FFT; same linear convolution)",
            "alt_one_long_fft": {"ms_per_step": ms_one, "value": out.numel() / (ms_one * 1e-3) / 1e9,
                                 "frac": alg / (ms_one * 1e-3) / 1e9 / cx.peak, "rel_l2_vs_float64_definition": err_one,
                                 "note": "FFTCONV_CUDA_CONV_SINGLE_FFT=1: one FFT of length 32768 per row pair "
                                         "(16 x 2048 through global scratch), the reference convolve_fftw's method"},
            "cpu_baseline": cpu_leg("conv", np.float64, 1024, 16384, 165) if cpu else None}


def config_c4(cx, steps=10, warmup=3, cpu=False):
    import numpy as np
    import scipy.signal
    import torch

    import fftconv_b200 as fc
    from fftconv_b200 import synthetic as syn

    dev = cx.dev
    x = torch.randn((8192, 8192), device=dev, generator=cx.g)
    env = torch.empty_like(x)
    ms = time_cuda(lambda: fc.hilbert_(x, env), steps, warmup)
    rows = [0, 4095, 8191]
    want = np.abs(scipy.signal.hilbert(x[rows].cpu().numpy().astype(np.float64), axis=-1))
    alg = 2 * 8192 * 8192 * 4
    return {"config": "c4", "workload": "batched Hilbert envelope f32 8192 x 8192",
            "ms_per_step": ms, "value": env.numel() / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "roofline": roofline(alg, ms, cx.peak), "algorithmic_bytes": alg, "l2": "inputs (268 MB) exceed L2",
            "rel_l2_vs_float64_definition": syn.rel_l2(env[rows].cpu().numpy(), want), "tolerance": 1e-5,
            "cpu_baseline": cpu_leg("hilbert", np.float32, 8192, 8192, 0) if cpu else None}


def config_c5(cx, log2n=31, steps=10, warmup=3, klen=245):
    """One long signal, K-1 tail exchanged between neighbouring ranks (NCCL) when world > 1.  The
    signal comes from the counter-based generator, so EVERY output can be checked from the
    definition on the host: the K outputs either side of every cut (the first K-1 after a cut are
    what the exchange produces), both ends (indices beyond 2^31 at the far end) and hashed-random
    interior points -- SURVEY.md section 8d."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from fftconv_b200 import multigpu as mg
    from fftconv_b200 import synthetic as syn

    dev, rank, world = cx.dev, cx.rank, cx.world
    n = 1 << log2n
    n_out = n + klen - 1
    lo, hi = mg.chunk_bounds(n, world, rank)
    a = torch.empty(hi - lo, device=dev)
    syn.fill_(a, syn.SEED_SIGNAL, first=lo)
    taps = syn.signal(syn.SEED_TAPS, 0, klen, np.float32, gain=1.0 / np.sqrt(klen))
    k = torch.from_numpy(taps).to(dev)

    def step():
        return mg.oaconvolve_long(a, k, rank=rank, world=world)

    for _ in range(max(warmup, 3)):
        y = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        y = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # ---- exact check of this rank's slice [lo, lo + len(y)) of the full result -------------------
    cuts = [mg.chunk_bounds(n, world, r)[0] for r in range(1, world)]
    # also the cuts a single GPU makes inside its chunk (CTA streams): probed with random points + ends
    idx = syn.cut_window_indices(cuts, n_out, klen, extra_random=6000, tail=4000, seed=rank)
    mine = idx[(idx >= lo) & (idx < lo + y.shape[0])]
    want = syn.conv_at(syn.SEED_SIGNAL, n, taps, mine)
    got = y[torch.from_numpy(mine - lo).to(dev)].cpu().numpy().astype(np.float64)
    num = torch.tensor([float(np.sum((got - want) ** 2)), float(np.sum(want ** 2)), float(mine.shape[0]),
                        float(np.sum(mine > 2 ** 31 - 1))], device=dev, dtype=torch.float64)
    # the first K-1 outputs after each cut, separately: these are sums of two ranks' contributions
    post = np.concatenate([np.arange(c, c + klen - 1) for c in cuts]) if cuts else np.zeros(0, np.int64)
    post = post[(post >= lo) & (post < lo + y.shape[0])]
    if post.shape[0]:
        w2 = syn.conv_at(syn.SEED_SIGNAL, n, taps, post)
        g2 = y[torch.from_numpy(post - lo).to(dev)].cpu().numpy().astype(np.float64)
        num2 = torch.tensor([float(np.sum((g2 - w2) ** 2)), float(np.sum(w2 ** 2)), float(post.shape[0])],
                            device=dev, dtype=torch.float64)
    else:
        num2 = torch.zeros(3, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(num, op=dist.ReduceOp.SUM)
        dist.all_reduce(num2, op=dist.ReduceOp.SUM)
    num, num2 = num.cpu().numpy(), num2.cpu().numpy()
    alg = (n + n_out) * 4 + klen * 4
    return {"config": "c5", "workload": f"single long signal f32 n=2^{log2n}, k={klen}, full, {world} GPU(s)",
            "n_gpus": world, "ms_per_step": ms, "value": n_out / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
            "roofline": roofline(alg, ms, cx.peak, world), "algorithmic_bytes": alg,
            "tail_exchange_bytes_per_cut": (klen - 1) * 4, "collective": "NCCL send/recv" if world > 1 else None,
            "data": "counter-based synthetic (fftconv_b200/synthetic.py), recomputed on the host for the check",
            "parity": {"rel_l2": float(np.sqrt(num[0] / max(num[1], 1e-300))), "checked_outputs": int(num[2]),
                       "checked_beyond_int32": int(num[3]),
                       "rel_l2_first_K-1_after_each_cut": (float(np.sqrt(num2[0] / max(num2[1], 1e-300)))
                                                           if num2[2] else None),
                       "exchanged_outputs_checked": int(num2[2]), "tolerance": 1e-5,
                       "what": "float64 direct sums at every output within K of each GPU cut and of both "
                               "signal ends, the last 4000 outputs, and 6000 hashed-random outputs per rank"}}


def config_rf(cx, steps=10, warmup=3):
    import torch

    import fftconv_b200 as fc

    dev = cx.dev
    rows, n, klen, dec = 8192, 8192, 245, 4
    x = torch.randn((rows, n), device=dev, generator=cx.g)
    k = torch.randn(klen, device=dev, generator=cx.g) / klen ** 0.5
    out = torch.empty((rows, n // dec), device=dev)
    kw = dict(log_compress=True, ref=1.0, dynamic_range_db=60.0, decimate=dec)
    ms_chain = time_cuda(lambda: fc.rf_envelope_(x, k, out, **kw), steps, warmup)
    ms_env = time_cuda(lambda: fc.envelope_(x, out, **kw), steps, warmup)
    filt, env = torch.empty_like(x), torch.empty_like(x)

    def separate():
        fc.oaconvolve_(x, k, filt, "same")
        fc.hilbert_(filt, env)
        return torch.clamp(1.0 + 20.0 * torch.log10(env[:, ::dec]) / 60.0, 0.0, 1.0)

    ms_sep = time_cuda(separate, steps, warmup)
    ms_hil = time_cuda(lambda: fc.hilbert_(x, env), steps, warmup)
    alg = rows * (n + n // dec) * 4 + klen * 4
    return {"config": "rf", "workload": f"RF chain f32 {rows} x {n}: band-pass k={klen} (same) -> envelope -> "
                                        f"log compression -> decimate by {dec}",
            "ms_per_step": ms_chain, "value": rows * n / (ms_chain * 1e-3) / 1e9, "unit": "Ginput-samples/s",
            "roofline": roofline(alg, ms_chain, cx.peak), "algorithmic_bytes": alg,
            "separate_passes_ms": ms_sep, "envelope_fused_ms": ms_env, "hilbert_plain_ms": ms_hil,
            "note": "separate = oaconvolve_ + hilbert_ + torch log10/clamp on a strided view"}


def config_pcie(cx):
    import torch

    dev = cx.dev
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
    return {"config": "pcie", "rank": cx.rank, "bytes_each_way": nbytes, "h2d_gbs": nbytes / t_h2d / 1e9,
            "d2h_gbs": nbytes / t_d2h / 1e9, "both_gbs_each_way": nbytes / t_both / 1e9,
            "c2_e2e_floor_ms": max(1073742804, 1077739520) / (nbytes / t_both) * 1e3}


def config_ksweep(cx, steps=6):
    import torch

    import fftconv_b200 as fc

    dev = cx.dev
    rows, n = 4096, 65536
    res = []
    for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
        r = rows if dt == torch.float32 else rows // 2
        a = torch.randn((r, n), device=dev, generator=cx.g, dtype=dt)
        for klen in (6, 20, 35, 65, 165, 245, 500, 768, 1000, 3000):
            k = torch.randn(klen, device=dev, generator=cx.g, dtype=dt)
            out = torch.empty((r, n + klen - 1), device=dev, dtype=dt)
            try:
                ms = time_cuda(lambda: fc.oaconvolve_(a, k, out, "full"), max(3, steps // 2), 3)
            except RuntimeError as e:
                res.append({"config": "ksweep", "dtype": name, "k": klen, "error": str(e)[:120]})
                continue
            alg = (a.numel() + out.numel() + klen) * a.element_size()
            res.append({"config": "ksweep", "dtype": name, "k": klen, "fft_size": fc.optimal_fft_size(klen), "rows": r,
                        "n": n, "ms_per_step": ms, "value": out.numel() / (ms * 1e-3) / 1e9, "unit": "Gsamples/s",
                        "frac_of_measured_hbm_peak": alg / (ms * 1e-3) / 1e9 / cx.peak})
        del a
    return res


def config_hostab(cx):
    """Host-pointer API on the headline shape: pinned buffers, and pageable buffers under the three
    strategies of FFTCONV_CUDA_HOST_PAGEABLE (0 driver staging, 1 cudaHostRegister per call, 2 ring)."""
    import numpy as np
    import torch

    import fftconv_b200 as fc

    rows, n, klen = 4096, 65536, 245
    rng = np.random.default_rng(1)
    a = rng.standard_normal((rows, n), dtype=np.float32)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(np.float32)
    out = np.zeros((rows, n + klen - 1), np.float32)

    def wall(fn, reps=3):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    res = {"config": "hostab", "workload": f"oaconvolve f32 {rows} x {n}, k={klen}, full, host pointers", "ms": {}}
    for strategy, name in (("0", "pageable_plain_memcpyasync"), ("1", "pageable_cudaHostRegister_per_call"),
                           ("2", "pageable_pinned_ring_3_threads")):
        os.environ["FFTCONV_CUDA_HOST_PAGEABLE"] = strategy
        res["ms"][name] = wall(lambda: fc.oaconvolve_(a, k, out, "full"))
    for mb in ("4", "8", "32"):
        os.environ["FFTCONV_CUDA_HOST_PAGEABLE"] = "2"
        os.environ["FFTCONV_CUDA_HOST_RING_BLOCK_MB"] = mb
        res["ms"][f"pageable_ring_block_{mb}MB"] = wall(lambda: fc.oaconvolve_(a, k, out, "full"))
    os.environ.pop("FFTCONV_CUDA_HOST_RING_BLOCK_MB", None)
    os.environ.pop("FFTCONV_CUDA_HOST_PAGEABLE", None)
    ap = torch.from_numpy(a).pin_memory()
    op = torch.empty(out.shape, dtype=torch.float32).pin_memory()
    res["ms"]["pinned"] = wall(lambda: fc.oaconvolve_(ap.numpy(), k, op.numpy(), "full"))
    t0 = time.perf_counter()
    np.copyto(out, op.numpy())
    res["ms"]["numpy_copy_1GB_single_thread"] = (time.perf_counter() - t0) * 1e3
    return res


def config_rowsweep(cx, steps=200):
    """The headline kernel on 1/8 ... 1/1 of the batch: what a rank sees under strong scaling."""
    import torch

    import fftconv_b200 as fc

    dev = cx.dev
    n, klen = 65536, 245
    a = torch.randn((4096, n), device=dev, generator=cx.g)
    k = torch.randn(klen, device=dev, generator=cx.g) / klen ** 0.5
    out = torch.empty((4096, n + klen - 1), device=dev)
    filt = fc.Filter(k)
    res = {"config": "rowsweep", "rows": {}}
    for rows in (512, 1024, 2048, 4096):
        ms = time_cuda(lambda: fc.oaconvolve_(a[:rows], k, out[:rows], "full"), steps, 5, cx.flush if rows < 4096 else None)
        msh = time_cuda(lambda: filt.oaconvolve_(a[:rows], out[:rows], "full"), steps, 5, cx.flush if rows < 4096 else None)
        res["rows"][rows] = {"ms_per_step": ms, "ms_per_step_filter_handle": msh}
    base = res["rows"][4096]["ms_per_step"]
    for rows, r in res["rows"].items():
        r["strong_scaling_efficiency_vs_4096"] = base * rows / 4096 / r["ms_per_step"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log2n", type=int, default=31, help="c5 signal length = 2**log2n")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cx = Ctx(dev, rank, world)
    which = ["c1", "c3", "c4", "c5"] if args.config == "all" else args.config.split(",")

    def emit(d):
        if rank == 0:
            for item in (d if isinstance(d, list) else [d]):
                print(json.dumps(item), flush=True)

    cpu = not args.no_cpu
    if "c1" in which and world == 1:
        emit(config_c1(cx, cpu=cpu))
    if "c3" in which and world == 1:
        emit(config_c3(cx, args.steps, args.warmup, cpu=cpu))
    if "c4" in which and world == 1:
        emit(config_c4(cx, args.steps, args.warmup, cpu=cpu))
    if "rf" in which and world == 1:
        emit(config_rf(cx, args.steps, args.warmup))
    if "pcie" in which:
        r = config_pcie(cx)
        if world > 1:  # every rank at once: does the host side keep up with N links?
            gathered = [None] * world
            dist.all_gather_object(gathered, r)
            r = gathered
        emit(r)
    if "ksweep" in which and world == 1:
        emit(config_ksweep(cx, args.steps))
    if "hostab" in which and world == 1:
        emit(config_hostab(cx))
    if "rowsweep" in which and world == 1:
        emit(config_rowsweep(cx))
    if "c5" in which:
        emit(config_c5(cx, args.log2n, args.steps, args.warmup))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
