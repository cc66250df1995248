"""BASELINE.json configs[3]: HugeRoom.pv scaled to 2048x2048, 8 listener positions ("sources") sharded across the GPUs of one
node, 4000 time steps each.  STRONG scaling: the 8 sources are fixed, every rank solves its contiguous shard in batches of at
most `--batch` sources (a source's pressure history is 67 GB at this size: two fit next to each other in 180 GB), then ONE
all-gather of the per-emitter outputs.  No data-path collective.

    python tools/gpu_config4.py                      # 1 GPU, 4 batches of 2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gpu_config4.py

Time = max over ranks of the device time (CUDA events of the solver stream around the whole shard) after one warm-up solve.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
from planeverb_b200 import pvcuda, sharding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--T", type=int, default=4000)
    ap.add_argument("--sources", type=int, default=8)
    ap.add_argument("--batch", type=int, default=0, help="sources per solve; 0 = as many as fit 95 %% of the device memory")
    ap.add_argument("--mem-gb", type=float, default=178.0, help="device memory the batch size is derived from")
    ap.add_argument("--scene", default="HugeRoom")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dist = tdev = None
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(local); tdev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=tdev)
    scenes = common.load_scenes()
    size, scale = common.scaled_config(a.n)
    boxes = common.boxes_of(scenes, a.scene, scale)
    everyone = common.listeners_for(a.sources, scale)
    lo, hi = sharding.shard_bounds(a.sources, world, rank)
    fit = sharding.max_batch_for_memory(lambda S: pvcuda.memory_requirement(a.n, a.n, a.T, S), 0.95 * a.mem_gb * 1e9, max(hi - lo, 1))
    if fit < 1:
        raise SystemExit(f"one {a.n}x{a.n} source of {a.T} steps does not fit {a.mem_gb} GB")
    plan = sharding.plan_batches(a.sources, world, rank, min(a.batch, fit) if a.batch > 0 else fit)
    B = max([h - l for l, h in plan] or [1])
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in common.EMITTERS]
    G = pvcuda.Scene(size, size, 275, T=a.T, max_sources=B, device=local if world > 1 else 0, efree=0.0447895788)
    assert G.gx == a.n and G.gy == a.n
    for b in boxes: G.add_aabb(*b)
    G.flush_geometry()
    batches = [everyone[l:h] for l, h in plan]
    bufs = [pvcuda.pinned_array((len(b), len(emitters), 8)) for b in batches]

    def sync():
        G.wait()
        if dist is not None:
            import torch
            dist.barrier(); torch.cuda.synchronize()

    if batches: G.solve_async(batches[0])          # warm-up (module load, first-touch of the history)
    sync()
    t0 = time.perf_counter()
    G.mark(0)
    tickets, step_ms, ana_ms = [], 0.0, 0.0
    for b, buf in zip(batches, bufs):
        G.solve_async(b)
        tickets.append(G.lookup_async(emitters, buf, n=len(b)))
    for t in tickets: G.lookup_wait(t)
    G.mark(1)
    G.wait()
    dev_ms = G.mark_elapsed_ms() if batches else 0.0
    st, an, _, _ = G.timing()                         # phase split of the last batch
    out = np.concatenate(bufs) if bufs else np.zeros((0, len(emitters), 8), np.float32)
    gathered = np.concatenate(sharding.gather_outputs(out, dist, tdev, n_total=a.sources))     # the path's one exchange
    sync()
    wall = time.perf_counter() - t0
    dev_s, wall_s, st_s, an_s = (float(v) for v in sharding.max_over_ranks([dev_ms / 1e3, wall, st / 1e3, an / 1e3], dist, tdev))
    if rank == 0:
        units = a.n * a.n * a.T * a.sources
        print(json.dumps({"config": f"{a.scene}.pv scaled to {a.n}x{a.n}, {a.sources} sources sharded over {world} GPU(s), {a.T} steps "
                                    f"(BASELINE.json configs[3]); batches of <= {B} sources per rank", "scaling": "strong", "n_gpus": world,
                          "device_ms": dev_s * 1e3, "wall_ms_incl_gather": wall_s * 1e3, "Mcell_updates_per_s": units / dev_s / 1e6,
                          "Mcell_updates_per_s_wall": units / wall_s / 1e6, "last_batch_ms": {"step_kernels": st_s * 1e3, "analyzer": an_s * 1e3},
                          "outputs_checksum": float(np.nan_to_num(gathered.astype(np.float64)).sum()), "outputs_shape": list(gathered.shape)}), flush=True)
    G.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
