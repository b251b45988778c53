"""Scene sharding logic, including a real 2-process gloo run (the N>1 path without GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pcops_b200.sharding import agree_on_region, max_over_ranks, shard_bounds, shard_sizes, sum_over_ranks


@pytest.mark.parametrize("n,world", [(312, 8), (16, 1), (16, 2), (7, 4), (3, 8), (0, 2), (1000003, 8)])
def test_shards_are_a_contiguous_balanced_partition(n, world):
    bounds = [shard_bounds(n, r, world) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a, b), (c, d) in zip(bounds, bounds[1:]):
        assert b == c and a <= b
    sizes = shard_sizes(n, world)
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def test_bad_rank_is_rejected():
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_units, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_units, rank, world)
    # each rank "processes" its scenes: a per-scene checksum that only depends on the scene id
    local = sum((s * 2654435761) % 1000003 for s in range(lo, hi))
    total = sum_over_ranks(local)
    slowest = max_over_ranks(1.0 + rank)          # rank-dependent "step time": everyone must see the max
    count = sum_over_ranks(hi - lo)
    if rank == 0:
        out.put((total, slowest, count))
    dist.destroy_process_group()


def test_two_rank_gloo_run_covers_every_scene_once():
    world, n_units = 2, 37
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r == 0 else None, n_units, q))
             for r in range(world)]
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    total, slowest, count = q.get()
    assert total == sum((s * 2654435761) % 1000003 for s in range(n_units))
    assert slowest == 2.0 and count == n_units


def _region_worker(rank, world, port, out):
    """An optional region that fails on rank 1 only: nobody may hang, everybody learns that it failed."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = None
    try:
        if rank == 1:
            raise ValueError("this rank's scan is rejected")
        local = 5.0 + rank
    except ValueError:
        pass
    slowest, ok = agree_on_region(local)
    slowest2, ok2 = agree_on_region(3.0 + rank)      # a second region in which every rank succeeds
    # a whole sweep in ONE collective: entry 1 failed on rank 1 only -> reported as failed, the others as the max
    from importlib import import_module
    vec = import_module("pcops_b200.sharding").max_vector_over_ranks([1.0 + rank, -1.0 if rank == 1 else 7.0, 2.0 - rank])
    assert vec == [2.0, -1.0, 2.0], vec
    if rank == 0:
        out.put((slowest, ok, slowest2, ok2))
    dist.destroy_process_group()


def test_a_rank_local_failure_in_an_optional_region_cannot_deadlock():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_region_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    slowest, ok, slowest2, ok2 = q.get()
    assert ok is False and slowest == 5.0
    assert ok2 is True and slowest2 == 4.0


def test_vector_reduce_single_process():
    from pcops_b200.sharding import max_vector_over_ranks
    assert max_vector_over_ranks([1.5, -1.0]) == [1.5, -1.0]
    assert max_vector_over_ranks([]) == []


def test_agree_on_region_single_process():
    assert agree_on_region(2.5) == (2.5, True)
    assert agree_on_region(None) == (-1.0, False)
