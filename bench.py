#!/usr/bin/env python
"""bench.py -- Planeverb hot path (2-D FDTD solve + per-cell analyzer) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

One "step" = one full pass of the hot path over one batch: Grid::GenerateResponse +
Analyzer::AnalyzeResponses (FDTD.cpp:87-236, Analyzer.cpp:48-431) for every listener position of the
batch.  Workload (BASELINE.json configs[2], the one north_star quotes its targets on): BigRoom.pv scaled
onto a 1024 x 1024 grid, 4 batched listener positions, 4000 time steps, resolution 275 -- per GPU.  With
N > 1 (torchrun, one rank per GPU) every rank solves its own 4 listener positions of the same scene
(weak scaling: independent sources shard with no data-path collective) and the ranks all-gather the
per-emitter acoustic parameters once per step over NCCL.

Metric: Mcell-updates/s = gx*gy*T*sources / seconds (interior cells x reference-equivalent steps).
  value     : inputs resident in HBM, timed with CUDA events on the solver stream, max over ranks
  e2e       : same, through the host-buffer C-ABI call (geometry + listeners uploaded, full result grids
              copied back to pinned host memory every step), wall clock bracketed by device syncs
  roofline  : the fused step kernel against the measured HBM copy bandwidth, 28 algorithmic bytes per
              cell-update
  cpu_baseline : the unmodified reference (oracle/_ref) on one host core over a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(scene="BigRoom", n=1024, T=4000, sources=4, resolution=275)
CPU_SAMPLE = dict(scene="BigRoom", n=1024, T=128, sources=1, resolution=275)
ALGO_BYTES_PER_CELL_UPDATE = 28          # r/w p,vx,vy (24) + one 4-byte wall coefficient
EFREE_275 = None                         # computed on the device at scene creation


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bench_listeners(k, scale):
    """k distinct, deterministic listener positions inside BigRoom's room (so every rank of a weak-scaling run
    gets statistically the same work): the Sandbox default (5, 0, 4) first, the rest on a fixed lattice walk."""
    out = []
    for i in range(k):
        x = 5.0 if i == 0 else 2.0 + (i * 1.37) % 6.0
        z = 4.0 if i == 0 else 2.0 + (i * 0.91) % 6.0
        out.append((x * scale, 0.0, z * scale))
    return out


def scene_inputs(cfg):
    from tests import common
    scenes = common.load_scenes()
    size, scale = common.scaled_config(cfg["n"], cfg["resolution"])
    return size, scale, common.boxes_of(scenes, cfg["scene"], scale), common


def run_reference_sample(steps, warmup, cfg=CPU_SAMPLE):
    """The reference's own CPU implementation of the path (oracle/_ref when the reference compiled in
    the build container, else the plain-C port) on the host cores of this box, single thread: the
    reference has no active parallel region (Analyzer.cpp:73,90 commented out, FDTD.cpp has none)."""
    size, scale, boxes, common = scene_inputs(cfg)
    listener = bench_listeners(1, scale)[0]
    from oracle import pvref
    kind = "reference"
    if pvref.available():
        sim = pvref.RefSim(size, size, cfg["resolution"], T=cfg["T"], efree=0.0447895788)
        for b in boxes:
            sim.add_aabb(*b)

        def one():
            sim.generate(listener)
            sim.analyze(listener)
    else:
        from oracle import pvoracle
        kind = "port"
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        sim = pvoracle.OracleSim(size, size, cfg["resolution"], T=cfg["T"], efree=0.0447895788)
        for b in boxes:
            sim.add_aabb(*b)

        def one():
            sim.generate(listener)
            sim.analyze(listener)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    units = cfg["n"] * cfg["n"] * cfg["T"] * cfg["sources"]
    return {"value": units / dt / 1e6, "unit": "Mcell-updates/s", "cores": 1, "kind": kind,
            "sample": f"{cfg['scene']}.pv scaled to {cfg['n']}x{cfg['n']}, {cfg['T']} steps, 1 listener, "
                      f"GenerateResponse+AnalyzeResponses, strict -O2 build, {steps} timed pass(es); host has {os.cpu_count()} cpus",
            "ms_per_step": dt * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=WORKLOAD["n"])
    ap.add_argument("--T", type=int, default=WORKLOAD["T"])
    ap.add_argument("--sources", type=int, default=WORKLOAD["sources"])
    ap.add_argument("--step-kernel", type=int, default=0)
    ap.add_argument("--variant", type=int, default=int(os.environ.get("PVC_VARIANT", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    peak, peak_src = load_peaks()
    cfg = dict(WORKLOAD, n=args.n, T=args.T, sources=args.sources)
    config = {"workload": f"{cfg['scene']}.pv scaled to {cfg['n']}x{cfg['n']} cells, {cfg['sources']} batched listener "
                          f"positions per GPU, {cfg['T']} time steps, resolution {cfg['resolution']} (BASELINE.json configs[2])",
              "grid": [cfg["n"], cfg["n"]], "time_steps": cfg["T"], "sources_per_gpu": cfg["sources"],
              "l2_policy": "inputs larger than L2: every step streams the pressure history "
                           f"({4 * cfg['n'] * cfg['n'] * cfg['T'] * cfg['sources'] / 1e9:.0f} GB) through HBM",
              "parallelism": f"sources sharded, {world} rank(s), no data-path collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference_sample(args.steps, args.warmup)
        line = {"impl": "reference", "metric": "Mcell-updates/sec (grid x steps / s), GenerateResponse+AnalyzeResponses",
                "value": r["value"], "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    from planeverb_b200 import pvcuda, sharding
    dist = None
    tdev = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        tdev = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", device_id=tdev)
    device = local_rank if world > 1 else 0

    size, scale, boxes, common = scene_inputs(cfg)
    S = cfg["sources"]
    listeners = sharding.shard(bench_listeners(S * world, scale), world, rank)      # S sources on every rank
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in common.EMITTERS]
    scene = pvcuda.Scene(size, size, cfg["resolution"], T=cfg["T"], max_sources=S, device=device,
                         step_kernel=args.step_kernel, variant=args.variant)
    assert scene.gx == cfg["n"] and scene.gy == cfg["n"]
    cells = scene.gx * scene.gy
    units_per_step = cells * scene.T * S                      # per rank

    def upload_geometry():
        scene.clear_geometry()
        for b in boxes:
            scene.add_aabb(*b)
        scene.flush_geometry()

    def sync_all():
        scene.wait()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    em_bufs = [pvcuda.pinned_array((S, len(emitters), 8)) for _ in range(2)]

    def exchange(out):
        """the one exchange of the path: per-emitter acoustic parameters of every source, all ranks (one NCCL all-gather)"""
        return np.concatenate(sharding.gather_outputs(out, dist, tdev, n_total=S * world))

    upload_geometry()
    # ---------------- device-resident timing (value) ----------------
    for _ in range(max(args.warmup, 3)):
        scene.solve_async(listeners)
    sync_all()
    sampler = ClockSampler(device)
    sampler.start()
    time.sleep(0.3)
    t_wall0 = time.perf_counter()
    scene.mark(0)
    launches, step_launches = 0, 0
    # frame loop, one frame deep: the emitter outputs of step k are copied out in stream order (lookup_async) and exchanged
    # between the ranks while step k+1 is already on the device
    gathered, ticket = None, None
    for k in range(args.steps):
        scene.solve_async(listeners)
        nl = scene.launch_counts()
        launches += nl[0] + nl[1]; step_launches += nl[0]
        t_new = scene.lookup_async(emitters, em_bufs[k & 1])
        if ticket is not None:
            scene.lookup_wait(ticket)
            gathered = exchange(em_bufs[(k - 1) & 1])
        ticket = t_new
    scene.lookup_wait(ticket)
    gathered = exchange(em_bufs[(args.steps - 1) & 1])
    scene.mark(1)                                  # after the last exchange: the timed region holds K solves and K exchanges
    st, an, _, _ = scene.timing()                  # phase split of the last timed solve (every solve runs the same launches)
    step_ms, ana_ms = st * args.steps, an * args.steps
    sync_all()
    t_wall1 = time.perf_counter()
    dev_ms = scene.mark_elapsed_ms()
    clocks = sampler.stop(t_wall0, t_wall1)

    # ---------------- end to end through the host-buffer C-ABI (e2e) ----------------
    # Frame loop as a plugin host drives it: every step uploads the geometry edit list and the listeners from host memory,
    # solves, and copies the full result grids into pinned host buffers (two sets, alternating).  The copy of step k runs
    # on the solver's copy stream under the time steps of step k+1 (pvx_solve_pipelined); the emitter outputs of step k are
    # read from its HOST grids and all-gathered.  Every step's grids are on the host before the clock stops.
    bufs = [(pvcuda.pinned_array((S, cells, 8)), pvcuda.pinned_array((S, cells))) for _ in range(2)]
    em_cells = [scene.emitter_cell(pos) for pos in emitters]

    def outputs_from_host(buf):
        out = np.full((S, len(emitters), 8), -1.0, np.float32)
        for e, rc_ in enumerate(em_cells):
            if rc_ is not None:
                out[:, e] = buf[0][:, rc_[0] * scene.gy + rc_[1]]
        return np.concatenate(sharding.gather_outputs(out, dist, tdev, n_total=S * world))

    upload_geometry(); scene.solve_pipelined(listeners, bufs[0]); scene.fetch_wait()          # warm
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        upload_geometry()
        scene.solve_pipelined(listeners, bufs[k & 1])          # returns once step k-1's grids are on the host
        if k > 0:
            outputs_from_host(bufs[(k - 1) & 1])
    scene.fetch_wait()
    gathered_e2e = outputs_from_host(bufs[(args.steps - 1) & 1])
    sync_all()
    e2e_s = time.perf_counter() - t0
    if not np.array_equal(np.nan_to_num(gathered_e2e), np.nan_to_num(gathered)):
        raise RuntimeError("bench: pipelined host-buffer outputs differ from the device-resident run")
    res, dly = bufs[0]
    h2d = len(boxes) * 24 + S * 24
    d2h = res.nbytes + dly.nbytes + S * len(emitters) * 32

    times = sharding.max_over_ranks([dev_ms / 1e3, e2e_s, step_ms / 1e3, ana_ms / 1e3], dist, tdev)
    dev_s, e2e_s, step_s, ana_s = (float(v) for v in times)
    total_units = units_per_step * world * args.steps

    if rank == 0:
        value = total_units / dev_s / 1e6
        step_gbs = ALGO_BYTES_PER_CELL_UPDATE * units_per_step * args.steps / step_s / 1e9
        traffic_path = os.path.join(ROOT, "profiles", "fused_step_traffic.json")
        per_gen = None
        if os.path.exists(traffic_path):
            try:
                per_gen = json.load(open(traffic_path)).get("dram_bytes_per_generation")
            except Exception:
                per_gen = None
        step_launches = max(step_launches, 1)        # step-kernel launches of the timed solves (the rest of gpu_launches: analyzer kernels)
        # ncu capture (profiles/): DRAM bytes of one generation (4 time steps, this grid, 4 sources), scaled to the
        # generations one launch of this run covers
        gens_per_launch = ((scene.T + 3) // 4) * args.steps / step_launches
        traffic = per_gen * gens_per_launch * (S / 4.0) if (os.path.exists(traffic_path) and per_gen and args.variant == 0 and args.step_kernel == 0) else None
        kernel_name = {0: "pvc::ws2::stepKernel<14,4,1,false,true> (default: warp-specialised generational kernel with producer + publisher warps, up to 256 x 4 time steps per launch)",
                       40: "pvc::ws2::stepKernel<15,4,1,false,false> (ws2 without the publisher warp)",
                       36: "pvc::fusedStepWsKernel<15,4,true> (first warp-specialised generational kernel)",
                       18: "pvc::fusedStepKernel<8,6,2> (one launch per 4 time steps)"}.get(args.variant, f"fused step kernel variant {args.variant}")
        if args.step_kernel == 1:
            kernel_name = "pvc::baselinePressureKernel + pvc::baselineVelocityKernel"
        line = {
            "metric": "Mcell-updates/sec (grid x steps / s), GenerateResponse+AnalyzeResponses",
            "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": total_units / e2e_s / 1e6, "unit": "Mcell-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kernel_name,
                         "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "launches_per_solve": step_launches / args.steps,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_UPDATE * units_per_step * args.steps / step_launches,
                         "avg_launch_us": step_s * 1e6 / step_launches,
                         "note": "achieved = 28 B per cell-update x cell-updates of the timed steps / CUDA-event time of the step-kernel phase; "
                                 "above 1.0 because 4 time steps are fused per tile pass (physical DRAM traffic: see traffic, bytes per launch)"},
            "phases_ms_per_step": {"step_kernels": step_s * 1e3 / args.steps, "analyzer": ana_s * 1e3 / args.steps},
            "wall_ms_per_step": (t_wall1 - t_wall0) * 1e3 / args.steps,
            "outputs_checksum": float(np.nan_to_num(gathered.astype(np.float64)).sum()),
        }
        if world == 1 and not args.no_cpu_baseline:
            r = run_reference_sample(1, 0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    scene.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
