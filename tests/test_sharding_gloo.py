"""N > 1 host logic on CPU: world_size-2 gloo run of the frame sharding and the timing reduction used by bench.py."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, PKG)
    from vslam_b200 import shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard.frame_range(rank, world, 2001))
    seeds = shard.weak_seeds(rank, 500)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, seeds))
    ms = shard.reduce_max(10.0 + 5.0 * rank, dist)          # slowest rank decides
    frames = shard.reduce_sum(len(mine), dist)
    dist.barrier()
    q.put((rank, gathered, ms, frames))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from vslam_b200 import shard

    for rank, gathered, ms, frames in res:
        all_frames = sorted(f for mine, _ in gathered for f in mine)
        assert all_frames == list(range(2001))                 # disjoint and complete
        seeds = [s for _, sd in gathered for s in sd]
        assert len(set(seeds)) == len(seeds) == 1000           # weak scaling: disjoint seeds per rank
        assert ms == 15.0 and frames == 2001
        assert shard.throughput(frames, ms) == pytest.approx(2001 / 0.015)


def test_frame_range_edge_cases():
    from vslam_b200 import shard

    assert list(shard.frame_range(0, 1, 5)) == [0, 1, 2, 3, 4]
    assert [len(shard.frame_range(r, 8, 3)) for r in range(8)] == [1, 1, 1, 0, 0, 0, 0, 0]
    assert sum(len(shard.frame_range(r, 4, 20000)) for r in range(4)) == 20000
    assert shard.weak_seeds(3, 4, unique=2) == [12, 13, 12, 13]
    # strong scaling: the ranks' seeds tile 0 .. total-1 exactly once whatever the world size
    for world in (1, 2, 3, 8):
        seeds = [s for r in range(world) for s in shard.strong_seeds(r, world, 20000)]
        assert seeds == list(range(20000))
    assert shard.strong_seeds(1, 2, 10, unique=2) == [5, 6, 5, 6, 5]
