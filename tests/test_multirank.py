"""CPU test of the N>1 path's host logic with world_size 2 over gloo: instance sharding covers every instance exactly
once, and the per-rank bank mixes reduce to the full mix on rank 0 (the same calls bench.py makes over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from klang_b200 import sharding


def test_shard_instances_partition():
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_instances(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert list(sharding.voice_ids(8, 128, 1, 2)) == list(range(512, 1024))
    with pytest.raises(ValueError):
        sharding.shard_instances(8, 2, 2)


def _instance_mix(i, n):
    """Stand-in for one instance's Synth::process output [2][n] (deterministic, exactly representable)."""
    t = np.arange(n, dtype=np.float32)
    return np.stack([(t % 7 - 3) * (i + 1), (t % 5 - 2) * (i + 1)]).astype(np.float32) * 0.25


def _worker(rank, world, port, total, n, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_instances(total, rank, world)
    mix = torch.zeros(2, n)
    for i in range(lo, hi):                      # per-rank bank mix in instance order (KB_BANK_MIX)
        mix += torch.from_numpy(_instance_mix(i, n))
    sharding.reduce_mix(mix, dst=0)
    if rank == 0:
        np.save(out, mix.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_mix_reduce(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total, n = 9, 256
    out = str(tmp_path / "mix.npy")
    mp.spawn(_worker, args=(2, port, total, n, out), nprocs=2, join=True)
    want = np.zeros((2, n), np.float32)
    for i in range(total):
        want += _instance_mix(i, n)
    assert np.array_equal(np.load(out), want)
