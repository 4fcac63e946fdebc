"""Multi-GPU partitioning of the hot path: one process per GPU, `torch.distributed`
(NCCL over NVLink on the box, gloo in CPU tests) for plumbing.

* Batches: rows are independent (the reference handles one signal per call,
  /root/reference/include/fftconv/fftconv.hpp:474-480) -> `shard_rows`, no collective.
* One long signal: overlap-add blocks are independent except for the K-1 sample
  output tail that block s adds into block s+1 (fftconv.hpp:387-410).  Cut the
  signal into one contiguous chunk per rank, convolve each chunk in Full mode, and
  hand the chunk's last K-1 outputs to the right neighbour, which adds them to its
  first K-1 outputs (<= 2 addends per sample: deterministic).  Message size
  4 (K-1) bytes -> latency bound: the tail is produced first by a tiny side-stream
  convolution of the chunk's last K-1 samples and sent while the main kernel runs.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_rows(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunk_bounds(n: int, world: int, rank: int, align: int = 4) -> Tuple[int, int]:
    """Contiguous sample range [lo, hi) of `rank` in a signal of n samples.  Interior
    cuts are multiples of `align` samples so every chunk starts 16-byte aligned
    (float32) when the whole signal does."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")

    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= world:
            return n
        return min(n, (n * r // world) // align * align)

    return cut(rank), cut(rank + 1)


_SIDE_STREAMS = {}


def _side_stream(device):
    """One exchange stream per device for the life of the process.  A fresh stream per call walks
    through torch's pool of 32, and the caching allocator keeps blocks per stream: the first ~40
    calls then each pay cudaMalloc for their small temporaries (measured at 8 GPUs: 1.2 ms per step
    falling to 0.6 ms)."""
    import torch

    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def oaconvolve_long(a_chunk, k, *, rank: int, world: int, group=None,
                    conv: Optional[Callable] = None, add: Optional[Callable] = None):
    """Full-mode overlap-add convolution of one long signal split across ranks.

    `a_chunk` is this rank's contiguous chunk (torch tensor, 1-D), `k` the taps
    (replicated).  Returns this rank's slice of the full result: chunk length
    samples, plus the final K-1 samples on the last rank.  `conv(a, k)` defaults to
    the CUDA `oaconvolve(..., "full")`, `add(dst, src)` to the CUDA in-place add;
    tests inject CPU stand-ins to exercise the exchange on gloo.
    """
    import torch
    import torch.distributed as dist

    own_conv = conv is None
    from . import api

    if conv is None or add is None:
        conv = conv or (lambda a, kk: api.oaconvolve(a, kk, "full"))
        add = add or api.add_
    klen = int(k.shape[0])
    m = int(a_chunk.shape[0])
    tail_len = klen - 1
    if world > 1 and m < tail_len:
        raise RuntimeError(f"chunk of {m} samples is shorter than the K-1 = {tail_len} tail")
    if world == 1 or tail_len == 0:
        return conv(a_chunk, k)

    # The tail (outputs m .. m+K-2 of the chunk's full convolution) depends only on the
    # chunk's last K-1 samples, so it is computed FIRST by a tiny convolution (taps and
    # samples swap roles: conv is commutative and the API wants signal >= kernel) on a
    # side stream and sent at once; the exchange then overlaps the main kernel.
    on_cuda = bool(getattr(a_chunk, "is_cuda", False))
    side = _side_stream(a_chunk.device) if on_cuda else None
    recv_buf = torch.empty(tail_len, dtype=a_chunk.dtype, device=a_chunk.device) if rank > 0 else None

    def exchange():
        ops = []
        send_buf = None
        if rank < world - 1:
            if own_conv and on_cuda:
                # the dedicated (K-1)^2 / 2 multiply-add kernel (fftconv_cuda_oa_tail)
                send_buf = torch.empty(tail_len, dtype=a_chunk.dtype, device=a_chunk.device)
                api.oa_tail_(a_chunk[m - tail_len:], k, send_buf)
            else:
                send_buf = conv(k, a_chunk[m - tail_len:].contiguous())[tail_len:].contiguous()
            ops.append(dist.P2POp(dist.isend, send_buf, rank + 1, group))
        if rank > 0:
            ops.append(dist.P2POp(dist.irecv, recv_buf, rank - 1, group))
        return dist.batch_isend_irecv(ops), send_buf

    if on_cuda:
        side.wait_stream(torch.cuda.current_stream(a_chunk.device))
        with torch.cuda.stream(side):
            reqs, send_buf = exchange()
    else:
        reqs, send_buf = exchange()
    # the main kernel is persistent (one CTA per SM): it leaves a few SMs to the exchange's kernel,
    # or the two wait for each other's SMs
    import os

    # measured in steady state (2^31 samples, K = 245): 2 ranks 2.47 / 2.26 / 2.31 ms with 0 / 1 / 4 SMs
    # left free (main kernel alone 2.18); 8 ranks 0.65-0.68 / 0.60 / 0.60-0.63 / 0.62-0.66 ms with
    # 0 / 2 / 4 / 8 (alone 0.57): interior ranks send and receive
    dflt = "1" if world == 2 else "2"
    reserve = int(os.environ.get("FFTCONV_B200_EXCHANGE_SMS", dflt)) if on_cuda and own_conv else 0
    if reserve:
        with api.reserved_sms(reserve):
            y = conv(a_chunk, k)              # m + K - 1 samples, main stream
    else:
        y = conv(a_chunk, k)
    for req in reqs:
        req.wait()                            # the current stream waits for the exchange
    if on_cuda:
        torch.cuda.current_stream(a_chunk.device).wait_stream(side)
        if send_buf is not None:
            send_buf.record_stream(torch.cuda.current_stream(a_chunk.device))
    if rank > 0:
        add(y[:tail_len], recv_buf)
    return y if rank == world - 1 else y[:m]
