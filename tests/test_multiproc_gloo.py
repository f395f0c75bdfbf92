"""World-size-2 checks of the multi-GPU host logic on CPU (gloo): sharding covers every unit exactly once,
job time is the max over ranks, and the MFCC all-gather reassembles the per-rank blocks in unit order.
The per-rank compute here is the ORACLE (CPU) -- the point is the plumbing, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jeicyboodsp_b200.sharding import allgather_features, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 4096, 36_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_utts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jeicyboodsp_b200 import synth
    from oracle.oracle import MfccParams, Oracle
    o = Oracle()
    p = MfccParams.preset("bench")
    b, e = shard_range(n_utts, rank, world)
    local = np.stack([o.mfcc_frames(synth.mfcc_utterance(u, 4000), p) for u in range(b, e)]).astype(np.float32)
    full = allgather_features(torch.from_numpy(local), n_utts)
    t = max_over_ranks(10.0 + rank)
    q.put((rank, full.numpy(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_and_max_time_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n_utts = 2, 5          # uneven shards: 3 + 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_utts, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from jeicyboodsp_b200 import synth
    from oracle.oracle import MfccParams, Oracle
    o = Oracle()
    expect = np.stack([o.mfcc_frames(synth.mfcc_utterance(u, 4000), MfccParams.preset("bench")) for u in range(n_utts)]).astype(np.float32)
    for rank, full, t in got:
        assert full.shape == expect.shape
        assert np.array_equal(full, expect)
        assert t == 11.0          # max over ranks of 10 + rank
