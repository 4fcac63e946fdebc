#!/usr/bin/env python
"""Headline benchmark: batched overlap-add convolution, float32, 4096 signals x 65536
samples, K = 245 taps, Full mode (BASELINE.json configs[1]) -- output Gsamples/s.

    python bench.py --gpus 1 --steps 20 --warmup 3            # our CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # reference CPU path
    torchrun ... bench.py --gpus N ...                        # N ranks, one per GPU

One "step" = one pass of the hot path over THE batch of 4096 signals.  Under torchrun the
fixed batch is sharded by rows (`multigpu.shard_rows`; rows are independent, no data-path
collective): STRONG scaling, as BASELINE.json's north star asks (">= 7x at 8 GPUs" on this
batch); the weak-scaling figure (4096 rows on every rank) rides along as `weak`.
`value` counts inputs resident in HBM; `e2e` is the same metric through the public
host-pointer API (pinned host buffers, H2D + compute + D2H inside the timed region);
`secondary` carries BASELINE.json's other configs (1, 3, 4, 5) with their parity figures.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "oaconvolve output Gsamples/s"
UNIT = "Gsamples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--k", type=int, default=245)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--soak-seconds", type=float, default=1.5,
                    help="untimed back-to-back steps before the timed ones: the part draws 1 kW and the clock "
                         "settles under the power cap only after about a second")
    ap.add_argument("--sustained-steps", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--c5-log2n", type=int, default=31)
    return ap.parse_args()


def workload_config(args):
    """Identical for both arms and every N: the workload, not the run."""
    return {
        "workload": f"batched oaconvolve float32, {args.rows} signals x {args.n} samples, k={args.k}, "
                    "mode=full (BASELINE.json configs[1])",
        "rows": args.rows, "n": args.n, "k": args.k, "mode": "full",
        "fft_size": "reference table: 2048 for k=245; large float32 jobs run 1024-point blocks (DESIGN.md 3.7)",
        "l2": "per-GPU inputs exceed the 126 MB L2 (1.07 GB per step at N = 1, 134 MB at N = 8); "
              "smaller shards get an L2 flush (256 MB write) between timed steps",
        "sharding": "the fixed batch is sharded by rows across ranks (strong scaling), no collective",
    }


# --------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons),
                "window": "soak + timed steps + sustained run (one continuous stretch of back-to-back kernels)"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy, read+write)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get("dram_bytes_per_launch")
    except Exception:
        return None


def bind_to_gpu_numa_node(local: int):
    """Run this rank's host side next to its GPU: CPU affinity (and so first-touch page placement of
    the pinned buffers) on the GPU's NUMA node.  Eight ranks that all sit on one socket push 8 x 2 GB
    of pinned traffic per step through that socket's memory controllers and the inter-socket link
    (round 1: e2e 25 -> 132 ms per step from 1 to 8 ranks)."""
    info = {"bound": False}
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node_path = f"/sys/bus/pci/devices/{bus}/numa_node"
        node = int(open(node_path).read().strip())
        info["gpu_pci"] = bus
        info["gpu_numa_node"] = node
        allowed = os.sched_getaffinity(0)
        info["allowed_cpus"] = len(allowed)
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        use = cpus & allowed
        info["node_cpus"] = len(cpus)
        if use:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["bound_cpus"] = len(use)
        else:
            info["note"] = "the GPU's NUMA node has no CPU in this process's cpuset"
    except Exception as e:  # best effort: never fail the bench over a missing sysfs file
        info["error"] = str(e)[:120]
    return info


# --------------------------------------------------------------------------------------
# CPU legs: the reference's own headers over the CPU FFT shim (oracle/_ref), all host
# threads, on a bounded sample of the same workload; second comparator: SciPy's pocketfft.
# --------------------------------------------------------------------------------------
def cpu_reference_rate(args, seconds, rows_cap=None):
    import numpy as np

    from oracle import ref

    if not ref.available():
        return None
    rng = np.random.default_rng(0x5EED0001)
    k = (rng.standard_normal(args.k) / np.sqrt(args.k)).astype(np.float32)
    threads = ref.hardware_threads()
    out_len = args.n + args.k - 1
    probe_rows = max(threads, 16)
    a = rng.standard_normal((probe_rows, args.n)).astype(np.float32)
    ref.oaconvolve(a[:threads], k, "full")  # plan creation / first touch
    t0 = time.perf_counter()
    ref.oaconvolve(a, k, "full")
    t_probe = time.perf_counter() - t0
    rows = int(max(probe_rows, min(args.rows, seconds / max(t_probe, 1e-6) * probe_rows)))
    if rows_cap:
        rows = min(rows, rows_cap)
    rows = max(threads, rows // threads * threads)
    a = rng.standard_normal((rows, args.n)).astype(np.float32)
    t0 = time.perf_counter()
    ref.oaconvolve(a, k, "full")
    dt = time.perf_counter() - t0
    return {"value": rows * out_len / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"first {rows} of {args.rows} signals x {args.n} samples, k={args.k}, full; "
                      f"{dt:.2f} s on {threads} std::thread workers; reference headers unmodified over "
                      "oracle/fftw_shim.cpp (FFTW itself is not installable here)",
            "_rows": rows, "_seconds": dt}


def cpu_scipy_rate(args, seconds):
    """A production CPU FFT on the same arrays: the shim under the
    reference headers is not FFTW, so its rate flatters any GPU/CPU ratio; pocketfft does not."""
    try:
        import numpy as np
        import scipy.fft
        import scipy.signal
    except Exception:
        return None
    rng = np.random.default_rng(0x5EED0001)
    k = (rng.standard_normal(args.k) / np.sqrt(args.k)).astype(np.float32)
    workers = len(os.sched_getaffinity(0))
    out_len = args.n + args.k - 1
    rows = max(16, workers)
    a = rng.standard_normal((rows, args.n)).astype(np.float32)

    def run(x):
        with scipy.fft.set_workers(workers):
            return scipy.signal.oaconvolve(x, k[None, :], mode="full", axes=-1)

    run(a[:2])
    t0 = time.perf_counter()
    run(a)
    tp = time.perf_counter() - t0
    rows2 = int(max(rows, min(args.rows, seconds / max(tp, 1e-6) * rows)))
    a = rng.standard_normal((rows2, args.n)).astype(np.float32)
    t0 = time.perf_counter()
    y = run(a)
    dt = time.perf_counter() - t0
    return {"value": rows2 * out_len / dt / 1e9, "unit": UNIT, "cores": workers, "kind": "scipy.signal.oaconvolve "
            "(pocketfft, scipy.fft.set_workers)", "dtype": str(y.dtype),
            "sample": f"first {rows2} of {args.rows} signals x {args.n} samples, k={args.k}, full; {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from oracle import ref

    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libfftconv_ref.so not built"}))
        return
    rng = np.random.default_rng(0x5EED0001)
    threads = ref.hardware_threads()
    k = (rng.standard_normal(args.k) / np.sqrt(args.k)).astype(np.float32)
    out_len = args.n + args.k - 1
    # size one step to ~3 s of CPU work
    probe = rng.standard_normal((max(threads, 16), args.n)).astype(np.float32)
    ref.oaconvolve(probe, k, "full")
    t0 = time.perf_counter()
    ref.oaconvolve(probe, k, "full")
    tp = time.perf_counter() - t0
    rows = int(max(probe.shape[0], min(args.rows, 3.0 / max(tp, 1e-6) * probe.shape[0])))
    rows = max(threads, rows // threads * threads)
    a = rng.standard_normal((rows, args.n)).astype(np.float32)
    for _ in range(args.warmup):
        ref.oaconvolve(a, k, "full")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.oaconvolve(a, k, "full")
    dt = (time.perf_counter() - t0) / args.steps
    value = rows * out_len / dt / 1e9
    sample = (f"each step = first {rows} of {args.rows} signals x {args.n} samples, k={args.k}, full, on "
              f"{threads} std::thread workers; reference headers unmodified over oracle/fftw_shim.cpp "
              "(a plain CPU FFT, about 5x slower per core than FFTW's codelets)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args), "sample_rows_per_step": rows,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                         "second_comparator": cpu_scipy_rate(args, 6.0)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
def timed_steps(fn, steps, barrier, flush=None):
    """K back-to-back steps between two events; with `flush`, an L2 flush before every step and
    only the steps themselves between events.  Returns total ms of the K steps on this rank."""
    import torch

    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    torch.cuda.synchronize()
    for a, b in ev:
        flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    barrier()
    return sum(a.elapsed_time(b) for a, b in ev)


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local)

    import torch
    import torch.distributed as dist

    import fftconv_b200 as fc
    from fftconv_b200 import multigpu as mg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the fixed batch, generated once with rank-independent content; every rank keeps its row shard
    lo, hi = mg.shard_rows(args.rows, world, rank)
    my_rows = hi - lo
    out_len = args.n + args.k - 1
    g = torch.Generator(device=dev).manual_seed(0x5EED0001)
    k = torch.randn(args.k, device=dev, generator=g) / args.k ** 0.5
    a_full = torch.randn((args.rows, args.n), device=dev, generator=g)       # weak-scaling leg uses all of it
    a = a_full[lo:hi]
    out_full = torch.empty((args.rows, out_len), device=dev)
    out = out_full[lo:hi]
    # L2 is 126 MB.  A rank's inputs are rows x 256 KiB: larger than L2 down to 512 rows per GPU (N = 8:
    # 134 MB in + 135 MB out per step); below that an L2 flush (256 MB write) runs between timed steps.
    need_flush = my_rows * args.n * 4 <= (126 << 20)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if need_flush else None
    flush = (lambda: flush_buf.zero_()) if need_flush else None

    def step():
        fc.oaconvolve_(a, k, out, "full")

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    # the first K steps on a cool part (before the soak): the burst figure, reported beside the steady state
    burst_ms = max_over_ranks(timed_steps(step, args.steps, barrier, flush)) / args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # soak: bring the part to its steady state under the power cap before timing
    t_end = time.perf_counter() + args.soak_seconds
    soak_steps = 0
    while time.perf_counter() < t_end:
        for _ in range(50):
            step()
        torch.cuda.synchronize()
        soak_steps += 50
    launches0 = fc.launch_count()
    ms_total = timed_steps(step, args.steps, barrier, flush)
    launches = fc.launch_count() - launches0
    ms_step = max_over_ranks(ms_total) / args.steps
    out_samples_total = args.rows * out_len
    value = out_samples_total / (ms_step * 1e-3) / 1e9

    # the same K-step measurement repeated over >= 1 s (the clock sampler sees the power-capped state)
    sus_steps = max(args.steps, args.sustained_steps if world == 1 else args.sustained_steps // 2)
    sus_ms = max_over_ranks(timed_steps(step, sus_steps, barrier, None)) / sus_steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- weak scaling, secondary: every rank convolves the whole 4096-row batch ----------------
    weak = None
    if world > 1:
        def step_weak():
            fc.oaconvolve_(a_full, k, out_full, "full")

        for _ in range(3):
            step_weak()
        wms = max_over_ranks(timed_steps(step_weak, max(10, args.steps), barrier, None)) / max(10, args.steps)
        weak = {"scaling": "weak", "rows_per_gpu": args.rows, "ms_per_step": wms,
                "value": world * out_samples_total / (wms * 1e-3) / 1e9, "unit": UNIT}

    # ---- the same step through a filter handle (taps transformed once) and a CUDA graph -----------
    handle = None
    try:
        filt = fc.Filter(k)

        def step_h():
            filt.oaconvolve_(a, out, "full")

        for _ in range(3):
            step_h()
        hms = max_over_ranks(timed_steps(step_h, max(20, args.steps), barrier, flush)) / max(20, args.steps)
        handle = {"ms_per_step": hms, "value": out_samples_total / (hms * 1e-3) / 1e9, "unit": UNIT,
                  "note": "fftconv_cuda_filter_oaconvolve: no spectrum kernel per step (SURVEY.md section 8f item 1)"}
        filt.close()
    except Exception as e:
        handle = {"error": str(e)[:200]}

    # ---- end to end through the host-pointer API -----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        a_h = torch.empty((my_rows, args.n), dtype=torch.float32).pin_memory()
        a_h.copy_(a)
        k_h = k.cpu()
        out_h = torch.empty((my_rows, out_len), dtype=torch.float32).pin_memory()
        an, kn, on = a_h.numpy(), k_h.numpy(), out_h.numpy()
        fc.oaconvolve_(an, kn, on, "full")  # warm the staging buffers and streams
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fc.oaconvolve_(an, kn, on, "full")
        dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        chk = float(torch.linalg.norm(out_h[:8].to(dev) - out[:8]) / torch.linalg.norm(out[:8]))
        e2e = {"value": out_samples_total / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": args.rows * args.n * 4 + world * args.k * 4,
               "d2h_bytes_per_step": args.rows * out_len * 4,
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "host_memory": "pinned (torch pin_memory), row blocks pipelined over 3 streams per GPU",
               "rel_l2_vs_device_path": chk, "numa": numa}
        del a_h, out_h
        # the reference API's real callers hand in std::vector / NumPy arrays: pageable memory
        try:
            ap = a.cpu().numpy()                       # plain malloc'ed arrays
            op = __import__("numpy").empty((my_rows, out_len), dtype="float32")
            op[:] = 0                                  # touch the pages
            fc.oaconvolve_(ap, kn, op, "full")
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                fc.oaconvolve_(ap, kn, op, "full")
            dtp = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
            chk_p = float(torch.linalg.norm(torch.from_numpy(op[:8]).to(dev) - out[:8]) / torch.linalg.norm(out[:8]))
            e2e["pageable"] = {"value": out_samples_total / dtp / 1e9, "unit": UNIT, "ms_per_step": dtp * 1e3,
                               "host_memory": "pageable numpy arrays through the library's page-locked ring "
                                              "(8 worker threads per GPU)",
                               "rel_l2_vs_device_path": chk_p}
            del ap, op
        except Exception as e:
            e2e["pageable"] = {"error": str(e)[:200]}

    # ---- BASELINE.json's other configs ---------------------------------------------------------------
    secondary = None
    if not args.no_secondary:
        import bench_configs as bcfg

        del a_full, out_full, a, out
        torch.cuda.empty_cache()
        cx = bcfg.Ctx(dev, rank, world)
        secondary = {}
        try:
            if world == 1:
                secondary["c1"] = bcfg.config_c1(cx)
                secondary["c3"] = bcfg.config_c3(cx)
                secondary["c4"] = bcfg.config_c4(cx)
            secondary["c5"] = bcfg.config_c5(cx, args.c5_log2n)
        except Exception as e:
            secondary["error"] = f"{type(e).__name__}: {str(e)[:300]}"

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes_total = args.rows * (args.n + out_len) * 4 + args.k * 4
        # rank 0's own shard: its kernel's algorithmic bytes over its own device time.  A step is the
        # spectrum kernel (2048 x K MACs, < 1 % at N = 1) followed by the segment kernel.
        own_ms = ms_total / args.steps
        alg_bytes = my_rows * (args.n + out_len) * 4 + args.k * 4
        achieved = alg_bytes / (own_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu_traffic() if world == 1 else None,
                    "kernel": "fc::w1k::oa_w1k_kernel (1024-point blocks, one warp per complex FFT)",
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": peak_src, "frac_of_8TBps_nominal": achieved / 8000.0,
                    "burst": {"steps": args.steps, "ms_per_step": burst_ms,
                              "frac": alg_bytes_total / world / (burst_ms * 1e-3) / 1e9 / peak,
                              "note": "the first K timed steps after warm-up, before the soak: the part has not reached "
                                      "its 1 kW power cap yet (SM clock near its maximum)"},
                    "sustained": {"steps": sus_steps, "ms_per_step": sus_ms,
                                  "frac": alg_bytes_total / world / (sus_ms * 1e-3) / 1e9 / peak}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_rate(args, args.cpu_seconds)
            if cpu:
                cpu = {kk: v for kk, v in cpu.items() if not kk.startswith("_")}
                cpu["second_comparator"] = cpu_scipy_rate(args, 8.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "soak_steps": soak_steps, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "rows_per_gpu": my_rows, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "weak": weak,
            "filter_handle": handle, "secondary": secondary,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
