"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): shard plans partition the job, the
single boundary all-gather reproduces the global signal inside every rank's local buffer, and the
shards' frame / tap coverage satisfies what the range entry points of the C ABI require."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyaudiorestoration_b200 import dist as pdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _signal(n, channels):
    rng = np.random.default_rng(7)
    return rng.standard_normal((channels, n)).astype(np.float32)


def _worker(rank, world, port, n, n_fft, hop, nt, channels, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = _signal(n, channels)
        sh = pdist.TimeShard(n, n_fft, hop, nt, rank, world)
        buf = sh.local_buffer(channels, "cpu")
        sh.chunk_view(buf).copy_(torch.from_numpy(x[:, sh.s0:sh.s1]))       # every rank only "has" its chunk
        sh.exchange_halos(buf)
        ok = np.array_equal(buf.numpy(), x[:, sh.origin:sh.origin + sh.local_len])
        curve = np.stack((np.linspace(0, 1, 50), 1 + 0.01 * np.arange(50)), -1) if rank == 0 else None
        got = pdist.broadcast_curve(curve, src=0)
        ok = ok and got.shape == (50, 2) and np.allclose(got[:, 1], 1 + 0.01 * np.arange(50))
        lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(lens, torch.tensor([sh.frame1 - sh.frame0]))
        q.put((rank, ok, sh.frame0, sh.frame1, [int(v) for v in lens]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    n, n_fft, hop, nt, channels = 50000, 1024, 256, 128, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, n_fft, hop, nt, channels, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get() for _ in range(world))
    assert all(r[1] for r in res)
    # frame ranges tile [0, T) in rank order
    T = n // hop + 1
    assert res[0][2] == 0 and res[-1][3] == T
    for a, b in zip(res, res[1:]):
        assert a[3] == b[2]
    assert sum(res[0][4]) == T


def test_plans_partition_and_cover():
    for n, n_fft, hop, nt, world in [(57600000, 4096, 1024, 128, 8), (100000, 512, 32, 50, 4), (186291, 4096, 1024, 8, 2),
                                     (5000, 64, 16, 20, 3)]:
        chunks = pdist.time_chunks(n, world, hop)
        assert chunks[0][0] == 0 and chunks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
        assert all(s0 % hop == 0 for s0, _ in chunks)
        frames = 0
        for r in range(world):
            sh = pdist.TimeShard(n, n_fft, hop, nt, r, world)
            frames += sh.frame1 - sh.frame0
            # what par_stft_range_f32 checks: the local slice covers the frames' samples (reflected at the ends)
            lo = sh.frame0 * hop - n_fft // 2
            hi = (sh.frame1 - 1) * hop + n_fft // 2
            assert max(lo, 0) >= sh.origin and min(hi, n) <= sh.origin + sh.local_len
            if hi > n:
                assert 2 * (n - 1) - (hi - 1) >= sh.origin
            # taps of outputs whose position rounds into [s0, s1]
            assert sh.s0 - nt >= sh.origin or sh.rank == 0
            assert sh.s1 + nt + 1 <= sh.origin + sh.local_len or sh.rank == world - 1
            assert sh.origin % 4 == 0
        assert frames == n // hop + 1
    assert list(pdist.shard_channels(8, 3, 8)) == [3]
    assert [len(pdist.shard_channels(7, r, 3)) for r in range(3)] == [3, 2, 2]
    assert sorted(c for r in range(3) for c in pdist.shard_channels(7, r, 3)) == list(range(7))
    with pytest.raises(ValueError):
        pdist.TimeShard(3000, 4096, 1024, 128, 0, 4)
