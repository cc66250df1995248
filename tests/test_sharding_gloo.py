"""CPU test of the N > 1 path (world_size 2, gloo): sharding the listener positions over ranks and
all-gathering the per-emitter outputs reproduces the single-process result.  The solve on each rank is
done by the oracle here (no GPU in this container) -- what is under test is the host-side sharding /
gather logic bench.py uses."""
import os
import socket

import numpy as np
import pytest

from planeverb_b200 import sharding
from tests import common


def test_shard_bounds_partition_everything():
    for n in (0, 1, 5, 8, 13):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [sharding.shard_bounds(n, world, r)[1] - sharding.shard_bounds(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _solve_with_oracle(listeners):
    from oracle import pvoracle
    scenes = common.load_scenes()
    out = np.zeros((len(listeners), len(common.EMITTERS), 8), np.float32)
    for i, L in enumerate(listeners):
        sim = pvoracle.OracleSim(25.0, 25.0, 275, T=80, efree=0.0447895788)
        for b in common.boxes_of(scenes, "SingleWall"):
            sim.add_aabb(*b)
        sim.generate(L)
        sim.analyze(L)
        for e, (x, z) in enumerate(common.EMITTERS):
            cell = int(np.float32(x) / sim.dx) * sim.gx + int(np.float32(z) / sim.dx)
            out[i, e] = sim.results[cell]
    return out


def _worker(rank, world, port, listeners, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard(listeners, world, rank)
    local = _solve_with_oracle(mine)
    gathered = sharding.gather_outputs(local, dist)
    # the frame loop's form: shard sizes derived from the list length, one collective
    known = sharding.gather_outputs(local, dist, n_total=len(listeners))
    assert [g.shape for g in known] == [g.shape for g in gathered] and all(np.array_equal(a, b) for a, b in zip(known, gathered))
    times = sharding.max_over_ranks([float(rank + 1), 0.5], dist)
    dist.barrier()
    q.put((rank, [g.copy() for g in gathered], times))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process():
    import torch.multiprocessing as mp
    listeners = common.listeners_for(3)            # 3 sources over 2 ranks: shards of 2 and 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, listeners, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _solve_with_oracle(listeners)
    for rank, gathered, times in got:
        assert [g.shape[0] for g in gathered] == [2, 1]
        merged = np.concatenate(gathered)
        assert np.array_equal(merged.view(np.uint32), want.view(np.uint32))
        assert list(times) == [2.0, 0.5]
