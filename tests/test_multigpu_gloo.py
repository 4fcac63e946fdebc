"""CPU, world size 2 over gloo: the host-side logic of the multi-GPU path (row
sharding, chunk bounds, K-1 tail exchange between neighbour ranks).  The CUDA
convolution is replaced by the oracle here -- only the partitioning and the
exchange protocol are under test; the same code runs on NCCL on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fftconv_b200 import multigpu as mg
from oracle import fftconv_oracle as O


def test_shard_rows_partitions_exactly():
    for batch in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [mg.shard_rows(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_chunk_bounds_cover_and_align():
    for n in (1000, 2 ** 20 + 3, 2 ** 31):
        for world in (1, 2, 4, 8):
            spans = [mg.chunk_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo % 4 == 0 for lo, _ in spans)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, klen, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(99)           # same signal on every rank
        a = rng.standard_normal(n).astype(np.float32)
        k = (rng.standard_normal(klen) / np.sqrt(klen)).astype(np.float32)
        lo, hi = mg.chunk_bounds(n, world, rank)

        def conv(x, kk):                          # CPU stand-in for the CUDA call
            return torch.from_numpy(O.oaconvolve(x.numpy(), kk.numpy(), "full"))

        def add(dst, src):
            dst += src

        y = mg.oaconvolve_long(torch.from_numpy(a[lo:hi].copy()), torch.from_numpy(k), rank=rank, world=world,
                               conv=conv, add=add)
        want = O.conv_exact(a, k, "full")
        end = hi if rank < world - 1 else n + klen - 1
        err = O.rel_l2(y.numpy(), want[lo:end])
        q.put((rank, int(y.shape[0]) == end - lo, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,klen", [(50_000, 245), (4001, 65), (9000, 300)])
def test_long_signal_tail_exchange_world2(n, klen):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, klen, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shape_ok, err in results:
        assert shape_ok, rank
        assert err < 1e-5, (rank, err)
