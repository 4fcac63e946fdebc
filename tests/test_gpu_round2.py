"""GPU: round-2 parity cases -- the host pipeline's concurrency (per-stream scratch, per-device
locks), pageable / tiny host paths, oversized Full outputs, the counter-based synthetic signal,
64-bit indexing beyond 2^31 samples, and the multi-GPU paths (torch.distributed + NCCL, and the
single-process C ABI `fftconv_cuda_mgpu_*`).  Tolerances are BASELINE.json's."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import fftconv_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t(x, dev):
    import torch

    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


# ---- the long-FFT scratch was shared by the host pipeline's three streams ------
@pytest.mark.parametrize("dt,n", [(np.float32, 32768), (np.float64, 16384), (np.float32, 20001)])
def test_host_hilbert_long_lines_many_blocks(cuda, dt, n, monkeypatch):
    """Host-pointer Hilbert of lines that take the split (global-scratch) FFT, with the row-block size
    forced down so that the batch runs as many blocks round-robin over the three pipeline streams."""
    import fftconv_b200 as fc

    monkeypatch.setenv("FFTCONV_CUDA_HOST_BLOCK_MB", "1")
    rng = np.random.default_rng(n)
    rows = 48 if n > 20001 else 64
    x = rng.standard_normal((rows, n)).astype(dt)
    for strategy in ("0", "2"):            # pinned-style pipeline on pageable memory, and the ring
        monkeypatch.setenv("FFTCONV_CUDA_HOST_PAGEABLE", strategy)
        monkeypatch.setenv("FFTCONV_CUDA_HOST_RING_BLOCK_MB", "1")
        got = fc.hilbert(x)
        assert O.rel_l2(got, O.hilbert_exact(x)) < H.TOL[np.dtype(dt)], strategy


def test_host_convolve_single_long_fft_many_blocks(cuda, monkeypatch):
    import fftconv_b200 as fc

    monkeypatch.setenv("FFTCONV_CUDA_HOST_BLOCK_MB", "1")
    monkeypatch.setenv("FFTCONV_CUDA_CONV_SINGLE_FFT", "1")
    rng = np.random.default_rng(5)
    a = rng.standard_normal((40, 30000))
    k = rng.standard_normal(165)
    got = fc.convolve(a, k, "full")
    assert O.rel_l2(got, O.conv_exact(a, k, "full")) < 1e-12


def test_two_streams_share_nothing(cuda):
    """Device pointers on two user streams at once: long-FFT Hilbert on both, results checked."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(11)
    xs = [rng.standard_normal((6, 65536)).astype(np.float32) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    outs = []
    torch.cuda.synchronize()
    for it in range(3):
        outs = []
        for x, st in zip(xs, streams):
            with torch.cuda.stream(st):
                xd = _t(x, cuda)
                outs.append(fc.hilbert(xd))
    torch.cuda.synchronize()
    for x, y in zip(xs, outs):
        assert O.rel_l2(y.cpu().numpy(), O.hilbert_exact(x)) < 1e-5
    for st in streams:
        fc.release_stream(st)              # frees that stream's workspace; later calls rebuild it
    with torch.cuda.stream(streams[0]):
        y = fc.hilbert(_t(xs[0], cuda))
    torch.cuda.synchronize()
    assert O.rel_l2(y.cpu().numpy(), O.hilbert_exact(xs[0])) < 1e-5


# ---- Full mode into an oversized output (fftconv.hpp:340, :385) -------------------------
@pytest.mark.parametrize("on_device", [False, True])
def test_full_mode_oversized_output_is_zero_filled(cuda, on_device):
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(3)
    a = rng.standard_normal((5, 3000)).astype(np.float32)
    k = rng.standard_normal(245).astype(np.float32)
    want = O.conv_exact(a, k, "full")
    if on_device:
        out = torch.full((5, 3300), float("nan"), device=cuda)
        fc.oaconvolve_(_t(a, cuda), _t(k, cuda), out[:, :3260], "full")   # 3260 > 3244, row stride 3300
        torch.cuda.synchronize()
        got = out.cpu().numpy()
    else:
        got = np.full((5, 3300), np.nan, np.float32)
        fc.oaconvolve_(a, k, got[:, :3260], "full")
    assert O.rel_l2(got[:, :3244], want) < 1e-5
    assert np.all(got[:, 3244:3260] == 0)
    assert np.all(np.isnan(got[:, 3260:]))
    with pytest.raises(RuntimeError, match="output length"):
        fc.oaconvolve_(a, k, np.empty((5, 3243), np.float32), "full")       # too short stays an error
    with pytest.raises(RuntimeError, match="output length"):
        fc.oaconvolve_(a, k, np.empty((5, 3001), np.float32), "same")       # Same stays exact


# ---- tiny host calls: one zero-copy kernel --------------------------------------------------------
def test_tiny_host_path_config1(cuda, monkeypatch):
    """BASELINE config 1 (n = 4352, K = 65) and its neighbours through host pointers: mapped
    page-locked buffer, one launch, no cudaMemcpy; must equal the device-pointer result bit for bit."""
    import torch

    import fftconv_b200 as fc

    monkeypatch.delenv("FFTCONV_CUDA_NO_SMALL_DIRECT", raising=False)
    rng = np.random.default_rng(21)
    for dt in (np.float32, np.float64):
        for batch, n, klen in ((1, 4352, 65), (1, 1664, 65), (3, 700, 120), (1, 300, 245), (2, 1000, 7), (1, 5, 3)):
            a = rng.standard_normal((batch, n)).astype(dt)
            k = rng.standard_normal(klen).astype(dt)
            for mode in ("full", "same"):
                before = fc.launch_count()
                got = fc.oaconvolve(a if batch > 1 else a[0], k, mode)
                assert fc.launch_count() - before == 1
                want = O.conv_exact(a if batch > 1 else a[0], k, mode)
                assert O.rel_l2(got, want) < H.TOL[np.dtype(dt)], (dt, batch, n, klen, mode)
                dev = fc.oaconvolve(_t(a if batch > 1 else a[0], cuda), _t(k, cuda), mode)
                torch.cuda.synchronize()
                assert np.array_equal(dev.cpu().numpy(), got), (dt, batch, n, klen, mode)
    monkeypatch.setenv("FFTCONV_CUDA_NO_TINY_HOST", "1")
    a, k = rng.standard_normal(4352).astype(np.float32), rng.standard_normal(65).astype(np.float32)
    assert O.rel_l2(fc.oaconvolve(a, k, "full"), O.conv_exact(a, k, "full")) < 1e-5


# ---- pageable host memory: three strategies, same samples ------------------------------------------
def test_pageable_host_strategies_agree(cuda, monkeypatch):
    import fftconv_b200 as fc

    rng = np.random.default_rng(31)
    a = rng.standard_normal((300, 20000)).astype(np.float32)        # 24 MB in, 24 MB out: above the 8 MB floor
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    want = O.conv_exact(a[:4], k, "full")
    outs = []
    for strategy in ("0", "1", "2"):
        monkeypatch.setenv("FFTCONV_CUDA_HOST_PAGEABLE", strategy)
        monkeypatch.setenv("FFTCONV_CUDA_HOST_RING_BLOCK_MB", "2")
        y = fc.oaconvolve(a, k, "full")
        assert O.rel_l2(y[:4], want) < 1e-5, strategy
        outs.append(y)
    # strided rows (views into wider arrays) through the ring
    wide_in = rng.standard_normal((300, 20100)).astype(np.float32)
    wide_out = np.full((300, 20300), np.nan, np.float32)
    fc.oaconvolve_(wide_in[:, 50:20050], k, wide_out[:, :20244], "full")
    assert O.rel_l2(wide_out[:3, :20244], O.conv_exact(wide_in[:3, 50:20050], k, "full")) < 1e-5
    assert np.all(np.isnan(wide_out[:, 20244:]))
    # the envelope options (thread-local call context) must reach the ring's worker threads
    x = rng.standard_normal((600, 8192)).astype(np.float32)
    e = fc.envelope(x, decimate=4, log_compress=True, ref=1.0, dynamic_range_db=50.0)
    assert O.rel_l2(e[:5], O.envelope(x[:5], decimate=4, log_compress=True, ref=1.0, dynamic_range_db=50.0)) < 1e-4
    assert O.rel_l2(e[-5:], O.envelope(x[-5:], decimate=4, log_compress=True, ref=1.0, dynamic_range_db=50.0)) < 1e-4


# ---- synthetic generator: device == numpy twin, bit for bit ------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_synthetic_fill_matches_numpy_twin(cuda, dt):
    import torch

    from fftconv_b200 import synthetic as syn

    tdt = torch.float32 if dt == np.float32 else torch.float64
    for first, n, gain in ((0, 100000, 1.0), (2 ** 33 + 12345, 4097, 0.25), (2 ** 31 - 50, 100, 1.0 / np.sqrt(245))):
        x = torch.empty(n, device=cuda, dtype=tdt)
        syn.fill_(x, syn.SEED_SIGNAL, first=first, gain=gain)
        torch.cuda.synchronize()
        assert np.array_equal(x.cpu().numpy(), syn.signal(syn.SEED_SIGNAL, first, n, dt, gain))
    big = syn.signal(7, 0, 1 << 20, np.float64)
    assert abs(big.mean()) < 5e-3 and abs(big.std() - 1.0) < 5e-3


def test_oa_tail_kernel(cuda):
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(41)
    for dt, klen in ((np.float32, 245), (np.float64, 165), (np.float32, 2)):
        k = rng.standard_normal(klen).astype(dt)
        last = rng.standard_normal(klen - 1).astype(dt)
        tail = torch.empty(klen - 1, device=cuda, dtype=torch.float32 if dt == np.float32 else torch.float64)
        fc.oa_tail_(_t(last, cuda), _t(k, cuda), tail)
        torch.cuda.synchronize()
        want = np.convolve(last.astype(np.float64), k.astype(np.float64))[klen - 1:]
        assert O.rel_l2(tail.cpu().numpy(), want) < H.TOL[np.dtype(dt)]


# ---- 64-bit indexing: one signal longer than 2^31 samples on ONE GPU (SURVEY section 7 hard part 5) ---
def test_long_signal_beyond_int32_indices(cuda):
    import torch

    import fftconv_b200 as fc
    from fftconv_b200 import synthetic as syn

    free, _ = torch.cuda.mem_get_info()
    n = 2 ** 31 + 4096 + 7
    if free < 3 * 4 * n:
        pytest.skip("needs ~26 GB of free device memory")
    klen = 245
    a = torch.empty(n, device=cuda)
    syn.fill_(a, syn.SEED_SIGNAL)
    taps = syn.signal(syn.SEED_TAPS, 0, klen, np.float32, gain=1.0 / np.sqrt(klen))
    for mode, n_out, pad in (("full", n + klen - 1, 0), ("same", n, klen // 2)):
        y = fc.oaconvolve(a, _t(taps, cuda), mode)
        torch.cuda.synchronize()
        assert y.shape[0] == n_out
        idx = syn.cut_window_indices([2 ** 31 - 1, 2 ** 31, 2 ** 30], n_out, klen, extra_random=20000, tail=6000, seed=1)
        want = syn.conv_at(syn.SEED_SIGNAL, n, taps, idx + pad)
        got = y[torch.from_numpy(idx).to(cuda)].cpu().numpy()
        assert np.sum(idx > 2 ** 31) > 4000
        assert syn.rel_l2(got, want) < 1e-5, mode
        # the far end on its own: every output index there exceeds INT32_MAX
        far = idx[idx > 2 ** 31]
        assert syn.rel_l2(y[torch.from_numpy(far).to(cuda)].cpu().numpy(), syn.conv_at(syn.SEED_SIGNAL, n, taps, far + pad)) < 1e-5
        del y
    del a
    torch.cuda.empty_cache()


# ---- multi-GPU ------------------------------------------------------------------------------------
def _visible_gpus():
    import torch

    return torch.cuda.device_count()


def test_long_signal_tail_exchange_over_nccl(cuda):
    """BASELINE config 5 end to end under torchrun + NCCL: every rank checks, from the definition, the
    outputs around ITS cut -- including the first K-1 after the cut, which are sums of its own
    partial result and the tail its left neighbour sent (fftconv.hpp:387-410 across the cut)."""
    g = _visible_gpus()
    if g < 2:
        pytest.skip("one GPU visible: the NCCL exchange needs two")
    world = min(g, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench_configs.py"),
           "--config", "c5", "--log2n", "27", "--steps", "3", "--gpus", str(world)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    p = d["parity"]
    assert d["n_gpus"] == world
    assert p["exchanged_outputs_checked"] == (world - 1) * 244
    assert p["rel_l2_first_K-1_after_each_cut"] < 1e-5, p
    assert p["rel_l2"] < 1e-5, p


def test_mgpu_c_abi(cuda):
    """fftconv_cuda_mgpu_*: one process over every visible GPU (one GPU: the same code without the
    exchange).  Batches by rows; one long signal host-to-host and device-resident."""
    import torch

    import fftconv_b200 as fc
    from fftconv_b200 import synthetic as syn

    g = _visible_gpus()
    mg = fc.MultiGpu()
    assert len(mg.devices) == g
    rng = np.random.default_rng(51)
    a = rng.standard_normal((37, 9000)).astype(np.float32)
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    for mode in ("full", "same"):
        assert O.rel_l2(mg.oaconvolve_batched(a, k, mode), O.conv_exact(a, k, mode)) < 1e-5
    x = rng.standard_normal((19, 4096))
    assert O.rel_l2(mg.hilbert_batched(x), O.hilbert_exact(x)) < 1e-12
    # long signal, host arrays, float64
    al = rng.standard_normal(3_000_000)
    kl = rng.standard_normal(165)
    yl = mg.oaconvolve_long_host(al, kl)
    cuts = [mg.chunk_bounds(al.shape[0], d)[0] for d in range(1, g)]
    idx = syn.cut_window_indices(cuts, yl.shape[0], 165, extra_random=3000, tail=1000)
    assert O.rel_l2(yl[idx], O.conv_exact_at(al, kl, idx)) < 1e-12
    # long signal, device resident, float32, synthetic: 2^26 samples
    n, klen = 1 << 26, 245
    taps = syn.signal(syn.SEED_TAPS, 0, klen, np.float32, gain=1.0 / np.sqrt(klen))
    ins, outs, los = [], [], []
    for d in range(g):
        lo, hi = mg.chunk_bounds(n, d)
        dev = torch.device("cuda", d)
        t = torch.empty(hi - lo, device=dev)
        syn.fill_(t, syn.SEED_SIGNAL, first=lo)
        ins.append(t)
        outs.append(torch.empty(hi - lo + klen - 1, device=dev))
        los.append(lo)
    for d in range(g):
        torch.cuda.synchronize(d)
    for _ in range(3):                     # repeated calls reuse the communicator's buffers
        mg.oaconvolve_long_(ins, taps, outs)
    mg.sync()
    n_out = n + klen - 1
    idx = syn.cut_window_indices(los[1:], n_out, klen, extra_random=5000, tail=2000)
    want = syn.conv_at(syn.SEED_SIGNAL, n, taps, idx)
    got = np.empty(idx.shape[0])
    for d in range(g):
        lo = los[d]
        hi = los[d + 1] if d + 1 < g else n_out
        m = (idx >= lo) & (idx < hi)
        got[m] = outs[d][torch.from_numpy(idx[m] - lo).to(outs[d].device)].cpu().numpy()
    assert syn.rel_l2(got, want) < 1e-5
    mg.close()


def test_two_threads_two_devices_do_not_serialise(cuda, monkeypatch):
    """Per-device locks: while one host thread is in the middle of a long
    host-pointer call on GPU 0 (tens of milliseconds of pipelined transfers), a second thread's
    small host-pointer call on GPU 1 must come back in its own time, not queue behind the first
    (round 1: one process-wide mutex held for the whole transfer).  Measured as latency, so that the
    two calls' competition for host memory bandwidth does not matter."""
    import time

    import torch

    import fftconv_b200 as fc

    if _visible_gpus() < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.delenv("FFTCONV_CUDA_NO_SMALL_DIRECT", raising=False)
    rng = np.random.default_rng(61)
    big = torch.from_numpy(rng.standard_normal((4096, 65536)).astype(np.float32)).pin_memory()
    big_out = torch.empty((4096, 65536 + 244)).pin_memory()
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    small = rng.standard_normal(4352).astype(np.float32)
    ks = rng.standard_normal(65).astype(np.float32)

    def long_call():
        with torch.cuda.device(0):
            fc.oaconvolve_(big.numpy(), k, big_out.numpy(), "full")

    def short_call():
        with torch.cuda.device(1):
            return fc.oaconvolve(small, ks, "full")

    long_call()                                  # warm both devices' pipelines
    short_call()
    t0 = time.perf_counter()
    long_call()
    t_long = time.perf_counter() - t0
    assert t_long > 0.010                         # the long call really is long
    th = threading.Thread(target=long_call)
    th.start()
    time.sleep(t_long * 0.2)                      # the long call is under way
    lat = []
    for _ in range(5):
        t0 = time.perf_counter()
        y = short_call()
        lat.append(time.perf_counter() - t0)
    still_running = th.is_alive()
    th.join()
    assert still_running, "the short calls did not overlap the long one"
    assert O.rel_l2(y, O.conv_exact(small, ks, "full")) < 1e-5
    assert O.rel_l2(big_out[:2].numpy(), O.conv_exact(big[:2].numpy(), k, "full")) < 1e-5
    assert min(lat) < 0.2 * t_long, (t_long, lat)


def test_threads_on_one_device_host_and_device_calls(cuda):
    """One device, three threads: two on device pointers (own streams), one on host pointers."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(71)
    a = rng.standard_normal((64, 30000)).astype(np.float32)
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    want = O.conv_exact(a[:2], k, "full")
    errors = []

    def dev_worker():
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                ad, kd = _t(a, cuda), _t(k, cuda)
                for _ in range(20):
                    y = fc.oaconvolve(ad, kd, "full")
                st.synchronize()
            assert O.rel_l2(y[:2].cpu().numpy(), want) < 1e-5
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def host_worker():
        try:
            for _ in range(4):
                y = fc.oaconvolve(a, k, "full")
            assert O.rel_l2(y[:2], want) < 1e-5
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ths = [threading.Thread(target=dev_worker), threading.Thread(target=dev_worker), threading.Thread(target=host_worker)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    assert not errors, errors


def test_filter_handle_device_mismatch_is_an_error(cuda):
    import torch

    import fftconv_b200 as fc

    if _visible_gpus() < 2:
        pytest.skip("needs two GPUs")
    k = np.ones(33, np.float32)
    f = fc.Filter(k)                        # lives on device 0
    a1 = torch.ones(1000, device="cuda:1")
    with pytest.raises(RuntimeError, match="device"):
        f.oaconvolve(a1, "full")
    f.close()


# ---- the warp-per-FFT engine (oa_w1k_*.cuh) -----------------------------------------------------------
W1K_CASES = [
    # batch, n, K
    (4, 65536, 245), (5, 10000, 245), (1, 100000, 245), (3, 5000, 300), (2, 780, 245), (1, 300, 245),
    (6, 7000, 400), (2, 4000, 8), (7, 1805, 65), (9, 3607, 256), (64, 65536, 245), (512, 65536, 245),
    (3, 400000, 165), (130, 9000, 246),
]


@pytest.mark.parametrize("batch,n,klen", W1K_CASES)
@pytest.mark.parametrize("mode", ["full", "same"])
def test_w1k_kernel_device_pointers(cuda, batch, n, klen, mode, monkeypatch):
    """1024-point blocks, one warp per complex FFT: row ends, odd batches, single rows, kernels from 8 to
    400 taps, both modes, against the definition and the oracle."""
    import torch

    import fftconv_b200 as fc

    monkeypatch.setenv("FFTCONV_CUDA_OA_FFT_SIZE", "1024")
    rng = np.random.default_rng(batch * 1000 + n + klen)
    a = rng.standard_normal((batch, n)).astype(np.float32)
    k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(np.float32)
    before = fc.launch_count()
    y = fc.oaconvolve(_t(a, cuda), _t(k, cuda), mode)
    torch.cuda.synchronize()
    assert fc.launch_count() - before == 1            # one launch: the kernel transforms the taps itself
    got = y.cpu().numpy()
    assert np.all(np.isfinite(got))
    pick = sorted(set([0, batch // 2, batch - 1]))
    assert O.rel_l2(got[pick], O.conv_exact(a[pick], k, mode)) < 1e-5
    if batch * n <= 1 << 22:
        assert O.rel_l2(got, O.oaconvolve(a, k, mode)) < 1e-5
    # the same samples as the 2048-point kernel produces (different blocks, same linear convolution)
    monkeypatch.setenv("FFTCONV_CUDA_NO_W1K", "1")
    monkeypatch.setenv("FFTCONV_CUDA_OA_FFT_SIZE", "2048")
    y2 = fc.oaconvolve(_t(a, cuda), _t(k, cuda), mode)
    torch.cuda.synchronize()
    assert O.rel_l2(got, y2.cpu().numpy()) < 1e-5


def test_w1k_kernel_alignment_and_guard_bands(cuda, monkeypatch):
    """Rows at every 4-byte alignment, padded row strides, NaN guard bands around every output row."""
    import torch

    import fftconv_b200 as fc

    monkeypatch.setenv("FFTCONV_CUDA_OA_FFT_SIZE", "1024")
    rng = np.random.default_rng(77)
    for off in range(4):
        for klen in (245, 246, 100):
            rows, n = 6, 9000 + off
            store_in = torch.full((rows * (n + 7) + 8,), float("nan"), device=cuda)
            a = rng.standard_normal((rows, n)).astype(np.float32)
            av = store_in[off:off + rows * (n + 7)].view(rows, n + 7)[:, :n]
            av.copy_(_t(a, cuda))
            k = (rng.standard_normal(klen) / 16).astype(np.float32)
            for mode in ("full", "same"):
                out_len = n + klen - 1 if mode == "full" else n
                ostore = torch.full((off + rows * (out_len + 5) + 8,), float("nan"), device=cuda)
                ov = ostore[off:off + rows * (out_len + 5)].view(rows, out_len + 5)
                fc.oaconvolve_(av, _t(k, cuda), ov[:, :out_len], mode)
                torch.cuda.synchronize()
                assert O.rel_l2(ov[:, :out_len].cpu().numpy(), O.conv_exact(a, k, mode)) < 1e-5, (off, klen, mode)
                assert torch.isnan(ov[:, out_len:]).all() and torch.isnan(ostore[:off]).all()


def test_w1k_is_the_default_for_large_float32_jobs(cuda, monkeypatch):
    """The throughput rule sends large float32 jobs with 8 .. 360 taps to 1024-point blocks; host pointers
    and the filter handle take the same path."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(78)
    a = rng.standard_normal((200, 65536)).astype(np.float32)
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    want = O.conv_exact(a[:3], k, "full")
    y = fc.oaconvolve(_t(a, cuda), _t(k, cuda), "full")
    monkeypatch.setenv("FFTCONV_CUDA_NO_W1K", "1")
    y2 = fc.oaconvolve(_t(a, cuda), _t(k, cuda), "full")
    monkeypatch.delenv("FFTCONV_CUDA_NO_W1K")
    torch.cuda.synchronize()
    assert O.rel_l2(y[:3].cpu().numpy(), want) < 1e-5 and O.rel_l2(y2[:3].cpu().numpy(), want) < 1e-5
    assert not torch.equal(y, y2)          # two different kernels did run
    assert O.rel_l2(fc.oaconvolve(a, k, "full")[:3], want) < 1e-5
    f = fc.Filter(k)
    assert O.rel_l2(f.oaconvolve(_t(a, cuda), "full")[:3].cpu().numpy(), want) < 1e-5
    f.close()


@pytest.mark.parametrize("klen", [17, 65, 245, 300])
def test_w1k_spectrum_in_kernel_matches_spectrum_kernel(cuda, klen, monkeypatch):
    """The warp-per-FFT kernel transforms the taps itself (one warp per CTA, the forward half of a unit, float32);
    FFTCONV_CUDA_W1K_SPECTRUM_KERNEL=1 restores the separate spectrum kernel (DFT in double).  Same answer
    to rounding, one launch instead of two, taps as host or device pointer."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(klen)
    a = rng.standard_normal((128, 40000)).astype(np.float32)       # > 4 Mi samples: the throughput rule applies
    k = (rng.standard_normal(klen) * np.exp(rng.uniform(-3, 3, klen))).astype(np.float32)   # taps of very different size
    ta, tk = _t(a, cuda), _t(k, cuda)
    before = fc.launch_count()
    y = fc.oaconvolve(ta, tk, "full")
    torch.cuda.synchronize()
    assert fc.launch_count() - before == 1
    y_host_taps = fc.oaconvolve(ta, k, "full")          # taps from host memory, signal on the device
    monkeypatch.setenv("FFTCONV_CUDA_W1K_SPECTRUM_KERNEL", "1")
    before = fc.launch_count()
    y2 = fc.oaconvolve(ta, tk, "full")
    torch.cuda.synchronize()
    assert fc.launch_count() - before == 2
    want = O.conv_exact(a[:4], k, "full")
    assert O.rel_l2(y[:4].cpu().numpy(), want) < 2e-6 and O.rel_l2(y2[:4].cpu().numpy(), want) < 2e-6
    assert O.rel_l2(y.cpu().numpy(), y2.cpu().numpy()) < 1e-6
    assert torch.equal(y, y_host_taps)


@pytest.mark.parametrize("pct", [15, 50, 90])
def test_w1k_dynamic_tail_is_bit_identical_and_rewinds_its_counters(cuda, pct, monkeypatch):
    """FFTCONV_CUDA_W1K_POOL_PCT=p: the last p per cent of a large job's units are handed out through an atomic
    counter instead of fixed ranges (oa_w1k_group.cuh: W1kParams::sched).  Which warp takes a unit does not change
    the unit's arithmetic, so the output is bit-identical to the fixed ranges -- on the first launch and on every
    later one (the kernel's last warp rewinds both counters), on a second stream too, for a job whose rows end in
    clipped blocks, in both modes."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(pct)
    a = rng.standard_normal((1200, 65536 - 24)).astype(np.float32)   # 50 units per warp: the dynamic rule applies
    k = (rng.standard_normal(245) / 16).astype(np.float32)
    ta, tk = _t(a, cuda), _t(k, cuda)
    for mode in ("full", "same"):
        want = fc.oaconvolve(ta, tk, mode)
        torch.cuda.synchronize()
        assert O.rel_l2(want[[0, 599, 1199]].cpu().numpy(), O.conv_exact(a[[0, 599, 1199]], k, mode)) < 1e-5
        monkeypatch.setenv("FFTCONV_CUDA_W1K_POOL_PCT", str(pct))
        for _ in range(3):
            y = fc.oaconvolve(ta, tk, mode)
            torch.cuda.synchronize()
            assert torch.equal(y, want), (mode, pct)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(2):
                y = fc.oaconvolve(ta, tk, mode)
        side.synchronize()
        assert torch.equal(y, want), (mode, pct, "second stream")
        monkeypatch.delenv("FFTCONV_CUDA_W1K_POOL_PCT")


def test_w1k_on_a_cut_down_grid(cuda):
    """tests/sanitize_w1k.py (written for compute-sanitizer, which this pool does not offer): the warp-per-FFT
    kernel on 6 CTAs -- fixed ranges and the dynamic tail, each launched twice, a filter handle, Same mode, clipped
    and misaligned rows -- against the float64 definition.  A grid of a few CTAs gives every warp long ranges and
    the pool many rounds, the opposite of the full-size runs."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "sanitize_w1k.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "sanitize w1k done" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---- the general 1-D FFT surface and Hilbert envelopes of any length -------------------
def _crel(y, ref) -> float:
    y, ref = np.asarray(y).astype(np.complex128).ravel(), np.asarray(ref).astype(np.complex128).ravel()
    d = np.linalg.norm(ref)
    return float(np.linalg.norm(y - ref) / (d if d > 0 else 1.0))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 8, 20, 64, 100, 243, 1000, 1024, 4096, 10007, 32768, 65536, 100000])
def test_general_fft_surface_matches_numpy(cuda, dt, n):
    """fftconv_cuda_fft_{c2c,r2c,c2r} (the reference's dft_1d / dft_r2c_1d / dft_c2r_1d wrappers,
    fftw.hpp:152-195) against numpy.fft in float64, host and device pointers, in place and out of place."""
    import torch

    import fftconv_b200 as fc

    tol = 2e-6 if dt == np.float32 else 1e-13
    cdt = np.complex64 if dt == np.float32 else np.complex128
    rng = np.random.default_rng(n)
    rows = 3
    z = (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))).astype(cdt)
    x = rng.standard_normal((rows, n)).astype(dt)
    z64, x64 = z.astype(np.complex128), x.astype(np.float64)
    assert _crel(fc.fft(z), np.fft.fft(z64)) < tol
    assert _crel(fc.fft(z, inverse=True), np.fft.ifft(z64) * n) < tol
    assert _crel(fc.fft(z[0]), np.fft.fft(z64[0])) < tol
    zt = torch.from_numpy(z).to(cuda)
    assert _crel(fc.fft(zt).cpu().numpy(), np.fft.fft(z64)) < tol
    X = fc.rfft(x)
    assert X.shape == (rows, n // 2 + 1) and X.dtype == cdt
    assert _crel(X, np.fft.rfft(x64)) < tol
    Xt = fc.rfft(torch.from_numpy(x).to(cuda))
    assert _crel(Xt.cpu().numpy(), np.fft.rfft(x64)) < tol
    back = fc.irfft(np.fft.rfft(x64).astype(cdt), n)
    assert back.dtype == dt and _crel(back, x64 * n) < tol
    backt = fc.irfft(Xt, n)
    assert _crel(backt.cpu().numpy(), x64 * n) < 2 * tol
    # in place through the C ABI (c2c only)
    from fftconv_b200 import _capi
    sfx = "f32" if dt == np.float32 else "f64"
    ip = zt.clone()
    v = torch.view_as_real(ip)
    _capi.check(getattr(_capi.lib(), f"fftconv_cuda_fft_c2c_{sfx}")(v.data_ptr(), v.data_ptr(), n, rows, -1, None))
    torch.cuda.synchronize()
    assert _crel(ip.cpu().numpy(), np.fft.fft(z64)) < tol


def test_general_fft_argument_errors(cuda):
    import fftconv_b200 as fc

    with pytest.raises(TypeError):
        fc.fft(np.zeros(8, np.float32))
    with pytest.raises(TypeError):
        fc.rfft(np.zeros(8, np.complex64))
    with pytest.raises(RuntimeError):
        fc.irfft(np.zeros(4, np.complex64), 8)
    with pytest.raises(RuntimeError):
        fc.fft(np.zeros((2, 0), np.complex64))


@pytest.mark.parametrize("dt,n", [(np.float32, 1 << 19), (np.float64, 1 << 18), (np.float32, 300001),
                                  (np.float64, 150001), (np.float32, 1 << 21)])
def test_hilbert_lines_beyond_the_split_fft(cuda, dt, n):
    """The reference takes any n (hilbert.hpp:16-22).  Lines longer than the on-chip and the split
    (16 x 16384) transforms -- which returned ERR_UNSUPPORTED in round 1 -- run on the general FFT."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(n & 0xffff)
    x = rng.standard_normal((3, n)).astype(dt)      # odd row count: the last pair is half empty
    want = O.hilbert_exact(x)
    got = fc.hilbert(x)
    assert _crel(got, want) < H.TOL[np.dtype(dt)]
    gt = fc.hilbert(torch.from_numpy(x).to(cuda))
    assert _crel(gt.cpu().numpy(), want) < H.TOL[np.dtype(dt)]
    # envelope post-processing rides the same path
    dec = fc.envelope(torch.from_numpy(x[:1]).to(cuda), decimate=4)
    assert _crel(dec.cpu().numpy(), want[:1, ::4]) < H.TOL[np.dtype(dt)]


# ---- the real-input Hilbert engine (hilbert_r2c_*.cuh): 8192-sample float32 lines ---------------------
@pytest.mark.parametrize("rows", [64, 67, 592, 1000, 2371])
def test_hilbert_r2c_engine_matches_oracle(cuda, rows, monkeypatch):
    """BASELINE config 4's line length on the real-input engine (one 4096-point complex FFT per line,
    mirror pass in registers) against scipy's float64 Hilbert transform and against the register-blocked
    engine it replaces; rows not a multiple of the groups per launch; NaN guard rows around the output."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(rows)
    x = rng.standard_normal((rows, 8192)).astype(np.float32)
    x[1] += 3.0                                   # a line with a DC offset (DC and Nyquist bins are zeroed)
    x[2, :] = 0.0
    x[2, 4097] = 1.0                              # an impulse
    xt = torch.from_numpy(x).to(cuda)
    out = torch.full((rows + 2, 8192), float("nan"), device=cuda)
    before = fc.launch_count()
    fc.hilbert_(xt, out[1:-1])
    torch.cuda.synchronize()
    assert fc.launch_count() - before == 1
    got = out.cpu().numpy()
    assert np.isnan(got[0]).all() and np.isnan(got[-1]).all()
    want = O.hilbert_exact(x)
    assert O.rel_l2(got[1:-1], want) < 1e-5
    assert np.abs(got[1:-1] - want).max() < 2e-5 * np.abs(want).max()
    monkeypatch.setenv("FFTCONV_CUDA_NO_HR2C", "1")
    old = fc.hilbert(xt).cpu().numpy()
    assert O.rel_l2(got[1:-1], old) < 1e-6


def test_hilbert_r2c_engine_needs_aligned_rows(cuda):
    """Rows that are not 16-byte aligned, strided rows and small batches take the other kernels; same answer."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(3)
    x = rng.standard_normal((70, 8192 + 8)).astype(np.float32)
    want = O.hilbert_exact(x[:, 1:8193])
    xt = torch.from_numpy(x).to(cuda)
    got = fc.hilbert(xt[:, 1:8193])               # misaligned rows, stride 8200
    assert O.rel_l2(got.cpu().numpy(), want) < 1e-5
    got4 = torch.empty((70, 8192), device=cuda)
    fc.hilbert_(xt[:, 4:8196], got4)              # aligned, strided: the r2c engine with a row stride
    assert O.rel_l2(got4.cpu().numpy(), O.hilbert_exact(x[:, 4:8196])) < 1e-5
    assert O.rel_l2(fc.hilbert(x[:5, :8192].copy()), O.hilbert_exact(x[:5, :8192])) < 1e-5   # small host batch


@pytest.mark.parametrize("dec,log", [(4, True), (8, False), (12, True), (3, False), (1, True)])
def test_hilbert_r2c_engine_envelope_epilogue(cuda, dec, log, monkeypatch):
    """Decimation / log compression in the store of the real-input engine (fftconv_cuda_envelope_*,
    SURVEY.md 8f-3) on 8192-sample lines, against the oracle restatement and the register-blocked kernel;
    and the rf_envelope chain (band-pass, envelope) on the same lines."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(100 + dec)
    rows = 130
    x = (rng.standard_normal((rows, 8192)) * np.exp(rng.uniform(-4, 1, (rows, 1)))).astype(np.float32)
    kw = dict(log_compress=log, ref=1.5, dynamic_range_db=50.0, decimate=dec)
    xt = torch.from_numpy(x).to(cuda)
    got = fc.envelope(xt, **kw).cpu().numpy()
    pick = [0, 1, 63, 64, 129]
    want = O.envelope(x[pick], **kw)
    assert got.shape == (rows, (8192 + dec - 1) // dec)
    if log:
        assert got.min() >= 0.0 and got.max() <= 1.0
        assert np.linalg.norm(got[pick].astype(np.float64) - want) / np.sqrt(want.size) < 4e-5
    else:
        assert O.rel_l2(got[pick], want) < 1e-5
    k = (rng.standard_normal(65) / 8).astype(np.float32)
    chain = fc.rf_envelope(xt, torch.from_numpy(k).to(cuda), **kw).cpu().numpy()
    monkeypatch.setenv("FFTCONV_CUDA_NO_HR2C", "1")
    old = fc.envelope(xt, **kw).cpu().numpy()
    old_chain = fc.rf_envelope(xt, torch.from_numpy(k).to(cuda), **kw).cpu().numpy()
    assert np.abs(got - old).max() < (2e-4 if log else 1e-5 * np.abs(old).max())
    assert np.abs(chain - old_chain).max() < (2e-4 if log else 1e-5 * np.abs(old_chain).max())


@pytest.mark.parametrize("rows,n", [(100, 8192), (700, 8192), (64, 4096), (600, 2048), (8, 1000), (4, 32768), (3, 300001)])
def test_hilbert_in_place(cuda, rows, n):
    """The reference's hilbert(x, env) works with env aliasing x (it reads x[i] right before it writes env[i],
    hilbert.hpp:57-63).  Every Hilbert kernel here must too: each reads a sample for the magnitude before the
    envelope is stored over it, and nothing reads a line after its envelope has been stored."""
    import torch

    import fftconv_b200 as fc

    rng = np.random.default_rng(rows + n)
    x = rng.standard_normal((rows, n)).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda)
    fc.hilbert_(xt, xt)
    torch.cuda.synchronize()
    pick = sorted(set([0, 1, rows // 2, rows - 1]))
    assert O.rel_l2(xt[pick].cpu().numpy(), O.hilbert_exact(x[pick])) < 1e-5
