"""world_size-2 gloo test of the multi-GPU plumbing (channel sharding, final gather, max-over-ranks).
The per-shard compute is the CPU oracle here (no GPU in this container); on the GPU box the same helpers
run under NCCL in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libtsd_b200.shard import channel_shard, gather_channels, max_over_ranks


def test_channel_shard_partitions():
    for nchan in (1, 7, 256, 1000, 4096):
        for world in (1, 2, 3, 8):
            spans = [channel_shard(nchan, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == nchan
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert channel_shard(256, 3, 8) == (96, 32)   # BASELINE config 4: 32 channels per GPU on 8 GPUs


def _worker(rank, world, port, nchan, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    P = oracle.port()
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((nchan, n)) + 1j * rng.standard_normal((nchan, n))).astype(np.complex64)
    h = P.design_rif_fen(31, "lp", 0.25)
    start, count = channel_shard(nchan, rank, world)
    y_local = np.stack([P.fir(1, h).step(x[c]) for c in range(start, start + count)]) if count else np.zeros((0, n), np.complex64)
    y = gather_channels(torch.from_numpy(y_local), nchan).numpy()
    t = max_over_ranks(10.0 + rank)
    if rank == 0:
        y_ref = np.stack([P.fir(1, h).step(x[c]) for c in range(nchan)])
        q.put((bool(np.array_equal(y, y_ref)), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nchan", [4, 5])
def test_two_rank_shard_and_gather(nchan):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nchan, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and t == 11.0
