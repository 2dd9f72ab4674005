"""world_size-2 gloo test of the clip-parallel path on CPU: the shards partition the clip set, the
merged per-clip results equal a single-process run bit for bit, and the timing reduction is a max."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eavsr_b200 import clip_parallel as CP
from eavsr_b200.synthetic import clip_inputs
from oracle import alignment as O

NUM_CLIPS = 5


def process_clip(i):
    """A CPU stand-in for one clip's forward: warp frame 1 onto frame 0 with a seeded flow."""
    clip = clip_inputs(1, 2, 16, 24, seed=100 + i)
    flow = torch.randn(1, 2, 16, 24, generator=torch.Generator().manual_seed(i)) * 2
    out = O.flow_warp(clip[:, 1], flow)
    return out.double().sum().item(), tuple(out.shape)


def _worker(rank, world_size, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert CP.world() == (rank, world_size, rank)
        local = CP.run_sharded(process_clip, NUM_CLIPS, rank, world_size)
        merged = CP.gather_results(local)
        slowest = CP.max_over_ranks(10.0 + rank)
        ret[rank] = (sorted(local), merged, slowest)
    finally:
        dist.destroy_process_group()


def test_two_rank_clip_parallel_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    single = {i: process_clip(i) for i in range(NUM_CLIPS)}
    assert ret[0][0] == [0, 2, 4] and ret[1][0] == [1, 3]
    for r in (0, 1):
        assert dict(ret[r][1]) == single            # bit-identical to the single-process run
        assert ret[r][2] == 11.0                    # max over ranks


def test_shard_properties():
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            shards = [CP.shard(n, r, w) for r in range(w)]
            assert sorted(i for s in shards for i in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        CP.shard(4, 2, 2)
    assert CP.max_over_ranks(3.5) == 3.5 and CP.gather_results({1: "a"}) == {1: "a"}
