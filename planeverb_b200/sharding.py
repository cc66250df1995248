"""Multi-GPU sharding of the hot path: the independent units are listener positions
("sources").  One process per GPU; every rank solves its own contiguous shard of the source list with no
data-path collective, and the ranks exchange only the per-emitter acoustic parameters (8 floats per
source x emitter) once at the end of a frame -- a single all-gather.

torch.distributed is plumbing here (NCCL over NVLink on the GPUs, gloo in the CPU tests); the solve itself
never sees a torch type.
"""
import numpy as np


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` of `world` (first ranks take the remainder)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items, world, rank):
    lo, hi = shard_bounds(len(items), world, rank)
    return list(items[lo:hi])


def plan_batches(n_items, world, rank, max_batch):
    """This rank's shard of n_items split into consecutive batches of at most max_batch items: [(lo, hi), ...] in global
    indices.  A 2048^2 source keeps 71 GB of pressure history for 4000 steps, so a GPU solves its shard of BASELINE
    configs[3] a few sources at a time (tools/gpu_config4.py); batches are as even as possible (5 items, max 2 -> 2, 2, 1)."""
    if max_batch < 1:
        raise ValueError(f"max_batch {max_batch}")
    lo, hi = shard_bounds(n_items, world, rank)
    n = hi - lo
    if n == 0:
        return []
    k = -(-n // max_batch)                      # number of batches
    base, extra = divmod(n, k)
    out, at = [], lo
    for b in range(k):
        size = base + (1 if b < extra else 0)
        out.append((at, at + size))
        at += size
    return out


def max_batch_for_memory(bytes_for, budget_bytes, cap):
    """Largest batch size S in 1..cap with bytes_for(S) <= budget_bytes (bytes_for = the solver's device-memory requirement
    for S batched sources, pvcuda.memory_requirement); 0 if not even one source fits."""
    best = 0
    for s in range(1, max(cap, 0) + 1):
        if bytes_for(s) <= budget_bytes:
            best = s
        else:
            break
    return best


def gather_outputs(local, dist=None, device=None, n_total=None):
    """All-gather per-source outputs.  local: float32 array [n_local_sources, n_emitters, 8].
    Returns the list of every rank's array, in rank order (shards may differ in length).
    n_total = length of the sharded source list: the shard sizes then follow from shard_bounds and ONE collective is
    issued (the frame loop's case); without it the sizes are exchanged first."""
    local = np.ascontiguousarray(local, np.float32)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    if n_total is not None:
        sizes = [hi - lo for lo, hi in (shard_bounds(n_total, world, r) for r in range(world))]
        if sizes[dist.get_rank()] != local.shape[0]:
            raise ValueError(f"rank {dist.get_rank()} holds {local.shape[0]} sources, its shard of {n_total} has {sizes[dist.get_rank()]}")
    else:
        count = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(count) for _ in range(world)]
        dist.all_gather(counts, count)
        sizes = [int(c) for c in torch.cat(counts).cpu().tolist()]
    most = max(sizes)
    padded = np.zeros((most,) + local.shape[1:], np.float32)
    padded[:local.shape[0]] = local
    mine = torch.from_numpy(padded).to(dev)
    everyone = torch.empty((world * most,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=dev)      # rank-major concatenation
    dist.all_gather_into_tensor(everyone, mine)
    host = everyone.cpu().numpy().reshape((world, most) + tuple(mine.shape[1:]))
    return [host[r, :sizes[r]] for r in range(world)]


def max_over_ranks(values, dist=None, device=None):
    """Element-wise max of a small float64 vector over all ranks (timing: the job takes as long as its slowest rank)."""
    values = np.asarray(values, np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return values
    import torch
    dev = device if device is not None else torch.device("cpu")
    t = torch.from_numpy(values.copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()
