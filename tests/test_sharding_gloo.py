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


def test_plan_batches_covers_the_shard_in_order():
    """Strong-scaling plan of BASELINE configs[3] (8 sources, at most 2 per solve): every source exactly once, rank-major,
    batches consecutive and as even as possible; more ranks than sources leaves the tail ranks empty."""
    for n, world, mb in [(8, 1, 2), (8, 2, 2), (8, 4, 2), (8, 8, 2), (5, 2, 2), (7, 3, 3), (3, 8, 2), (9, 2, 4)]:
        seen = []
        for r in range(world):
            plan = sharding.plan_batches(n, world, r, mb)
            lo, hi = sharding.shard_bounds(n, world, r)
            assert [b for l, h in plan for b in range(l, h)] == list(range(lo, hi))
            sizes = [h - l for l, h in plan]
            assert all(1 <= s <= mb for s in sizes) and (not sizes or max(sizes) - min(sizes) <= 1)
            assert len(plan) == -(-(hi - lo) // mb)
            seen += [b for l, h in plan for b in range(l, h)]
        assert seen == list(range(n))
    assert sharding.plan_batches(5, 1, 0, 2) == [(0, 2), (2, 4), (4, 5)]
    with pytest.raises(ValueError):
        sharding.plan_batches(4, 1, 0, 0)


def test_batch_size_from_the_memory_requirement():
    """pvc_memory_requirement is host arithmetic (no GPU): a 2048^2 source with 4000 steps of history needs ~72 GB, so two fit
    a 180 GB B200 and three do not; at 1024^2 nine do."""
    from planeverb_b200 import pvcuda
    one = pvcuda.memory_requirement(2048, 2048, 4000, 1)
    assert 70e9 < one < 74e9
    assert pvcuda.memory_requirement(2048, 2048, 4000, 2) > 2 * one * 0.99
    budget = 0.95 * 178e9
    assert sharding.max_batch_for_memory(lambda s: pvcuda.memory_requirement(2048, 2048, 4000, s), budget, 8) == 2
    assert sharding.max_batch_for_memory(lambda s: pvcuda.memory_requirement(1024, 1024, 4000, s), budget, 16) == 8
    assert sharding.max_batch_for_memory(lambda s: pvcuda.memory_requirement(2048, 2048, 4000, s), 10e9, 8) == 0
    assert pvcuda.memory_requirement(1, 1, 10, 1) == 0                      # invalid config


def test_streamed_solver_memory_requirement():
    """pvc_memory_requirement_streamed (host arithmetic): the history shrinks with history_steps, the carry planes and the K - 2
    state checkpoints are small beside it -- all eight 2048^2 x 4000-step sources of BASELINE configs[3] fit ONE B200 with an
    800-sample history, and a 4096^2 source (268 GB of full history) fits with 1600."""
    from planeverb_b200 import pvcuda
    full = pvcuda.memory_requirement(2048, 2048, 4000, 8)
    assert full > 560e9
    need = [pvcuda.memory_requirement(2048, 2048, 4000, 8, history_steps=h) for h in (400, 800, 1000, 2000, 4000)]
    assert all(a < b for a, b in zip(need, need[1:]))
    assert need[1] < 0.9 * 178e9 < need[3]
    assert need[4] > full                                               # one chunk: the full history plus the carry planes
    assert pvcuda.memory_requirement(2048, 2048, 4000, 8, history_steps=803) == need[1]      # rounded down to a multiple of 8
    assert pvcuda.memory_requirement(4096, 4096, 4000, 1) > 250e9
    assert pvcuda.memory_requirement(4096, 4096, 4000, 1, history_steps=1600) < 0.9 * 178e9


def _batched_worker(rank, world, port, listeners, max_batch, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    outs = [_solve_with_oracle(listeners[lo:hi]) for lo, hi in sharding.plan_batches(len(listeners), world, rank, max_batch)]
    local = np.concatenate(outs) if outs else np.zeros((0, len(common.EMITTERS), 8), np.float32)
    gathered = sharding.gather_outputs(local, dist, n_total=len(listeners))
    dist.barrier()
    q.put((rank, np.concatenate(gathered)))
    dist.destroy_process_group()


def test_two_rank_gloo_batched_strong_scaling_matches_single_process():
    """tools/gpu_config4.py's host logic on CPU: 5 fixed sources over 2 ranks in batches of at most 2 (plans 2+1 and 2), the
    oracle standing in for the device, one all-gather at the end -- the merged table equals the single-process one."""
    import torch.multiprocessing as mp
    listeners = common.listeners_for(5)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_batched_worker, args=(r, 2, port, listeners, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _solve_with_oracle(listeners)
    for rank, merged in got:
        assert merged.shape == want.shape and np.array_equal(merged.view(np.uint32), want.view(np.uint32))
