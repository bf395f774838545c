"""CPU-only: the N>1 path's host logic under a real 2-process gloo group (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ensemble_svs_with_interactions_b200 import sharding


def test_assign_is_a_balanced_partition():
    lengths = [6000, 5900, 400, 3000, 2800, 2500, 1200, 1100, 900, 50, 6000, 10]
    for world in (1, 2, 4, 8):
        parts = sharding.assign(lengths, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lengths)))                       # disjoint cover
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)                 # LPT bound
        for p in parts:
            assert [lengths[i] for i in p] == sorted((lengths[i] for i in p), reverse=True)
    assert sharding.assign([], 4) == [[], [], [], []]                  # empty input
    assert sharding.assign([5], 3) == [[0], [], []]                    # fewer items than ranks


def test_batches_respect_frame_budget():
    lengths = [2000, 1900, 1800, 900, 800, 100]
    bs = sharding.batches(list(range(6)), lengths, max_frames=4000)
    assert [i for b in bs for i in b] == list(range(6))
    for b in bs:
        assert len(b) * max(lengths[i] for i in b) <= 4000 or len(b) == 1
    assert sharding.batches([], lengths, 100) == []
    assert sharding.batches([0], lengths, 10) == [[0]]                 # an over-budget item still gets its own batch


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [100 * (i % 7 + 1) for i in range(23)]
    mine = sharding.assign(lengths, world)[rank]
    # every rank computes the same partition; gather the shards to prove it is a disjoint cover across PROCESSES
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = sharding.max_over_ranks(1.0 + rank)                            # slowest rank defines the clock
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, gathered, t))


def test_two_process_gloo_sharding_and_clock():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, t in res:
        assert sorted(i for part in gathered for i in part) == list(range(23))
        assert set(gathered[0]).isdisjoint(gathered[1])
        assert t == 2.0
