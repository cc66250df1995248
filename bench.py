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
  roofline  : the step kernel against the measured HBM copy bandwidth, 28 algorithmic bytes per cell-update
               -- plus the two roofs that actually bind it: physical DRAM bytes and issue slots
              (from the committed ncu capture of the same kernel, profiles/)
  verified  : (N = 1, default) the emitter outputs of the TIMED run, all sources, compared with the oracle
              (oracle/pv_oracle.c, pinned to the unmodified reference) outside the timed region
  cpu_baseline : the unmodified reference (oracle/_ref) on one host core over a bounded sample, strict build
              and (fast_math) the shipped Release flags
  extras    : the other BASELINE configs through the same library: the Sandbox contract case and configs[1]
              (latency), configs[3] (fixed 8-source 2048^2 job sharded over the N GPUs: strong scaling),
              configs[4] (dynamic-geometry frame loop, 8 sources over the N GPUs); at N = 1 also configs[3] as ONE
              batch of 8 on the streamed solver (bounded history, chunks recomputed for the backward pass)

The product arm imports nothing from oracle/ or tests/ except inside verify() and the cpu_baseline leg; every
PVC_* environment variable is removed at start-up (the release library reads none anyway).
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
EMITTERS = [(5, 6), (6, 5), (3.5, 3.5), (12.5, 12.5), (20, 20)]          # SURVEY.md 8d, pre-scale metres
ALGO_BYTES_PER_CELL_UPDATE = 28          # r/w p,vx,vy (24) + one 4-byte wall coefficient
EFREE_275 = 0.0447895788                 # FreeGrid at resolution 275; the device computes its own at scene creation
KERNEL_NCU = os.path.join(ROOT, "profiles", "step_kernel_ncu.json")
METRIC = "Mcell-updates/sec (grid x steps / s), GenerateResponse+AnalyzeResponses"


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
    """(size_m, scale, boxes) of a BASELINE-style scaled scene, through the product's own host derivation (pvx_derive)."""
    from planeverb_b200 import scenes as pscenes
    size, scale = pscenes.scaled_config(cfg["n"], cfg["resolution"])
    return size, scale, pscenes.boxes_of(pscenes.load_scenes(), cfg["scene"], scale)


def workload_text(cfg, per="per GPU"):
    return (f"{cfg['scene']}.pv scaled to {cfg['n']}x{cfg['n']} cells, {cfg['sources']} batched listener position(s) {per}, "
            f"{cfg['T']} time steps, resolution {cfg['resolution']}")


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation (the ONLY place besides verify() that touches oracle/)
# ------------------------------------------------------------------------------------------------------------------
def run_reference_sample(steps, warmup, cfg=CPU_SAMPLE):
    """The reference's own CPU implementation of the path (oracle/_ref when the reference compiled in
    the build container, else the plain-C port) on the host cores of this box, single thread: the
    reference has no active parallel region (Analyzer.cpp:73,90 commented out, FDTD.cpp has none)."""
    size, scale, boxes = scene_inputs(cfg)
    listener = bench_listeners(1, scale)[0]
    from oracle import pvref
    kind = "reference"
    if pvref.available():
        sim = pvref.RefSim(size, size, cfg["resolution"], T=cfg["T"], efree=EFREE_275)
    else:
        from oracle import pvoracle
        kind = "port"
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        sim = pvoracle.OracleSim(size, size, cfg["resolution"], T=cfg["T"], efree=EFREE_275)
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
    build = "shipped Release flags (-O3 -ffast-math -mavx2 -mfma ~ /O2 /fp:fast /arch:AVX2)" if os.environ.get("PVREF_LIB") else "strict -O2 -ffp-contract=off build"
    return {"value": units / dt / 1e6, "unit": "Mcell-updates/s", "cores": 1, "kind": kind,
            "sample": f"{workload_text(cfg, 'in all')}, GenerateResponse+AnalyzeResponses, {build}, {steps} timed pass(es) after {warmup} warm-up; "
                      f"1 thread (the reference has no active parallel region); host has {os.cpu_count()} cpus",
            "ms_per_step": dt * 1e3}


def fast_math_row(steps=1, warmup=0):
    """BASELINE.md section 3's second CPU row: the same unmodified sources built with the shipped Release flags, timed in a child
    process (the two builds export the same symbols).  None if that build is absent or this CPU lacks AVX2/FMA."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libpvref_fast.so")
    try:
        flags = open("/proc/cpuinfo").read()
    except Exception:
        flags = ""
    if not os.path.exists(lib) or " avx2" not in flags or " fma" not in flags:
        return None
    env = dict(os.environ, PVREF_LIB=lib)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup), "--no-fast-row"],
                             env=env, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
        line = json.loads(out[-1])
        return {"value": line["value"], "unit": line["unit"], "cores": 1, "flags": "-O3 -ffast-math -mavx2 -mfma (the shipped /O2 /fp:fast /arch:AVX2)"}
    except Exception:
        return None


def reference_line(args):
    r = run_reference_sample(args.steps, args.warmup)
    cfg = CPU_SAMPLE
    config = {"workload": workload_text(cfg, "in all") + " -- a BOUNDED SAMPLE of BASELINE.json configs[2] (same scene and grid, 128 of its 4000 steps, 1 of its "
                          "4 listeners); the metric is a rate, and a short run flatters the CPU (few cells have an onset yet, so the analyzer is nearly free)",
              "grid": [cfg["n"], cfg["n"]], "time_steps": cfg["T"], "sources": cfg["sources"],
              "gpu_arm_workload": workload_text(WORKLOAD)}
    cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if not args.no_fast_row:
        fm = fast_math_row(max(1, min(args.steps, 2)), 0)
        if fm:
            cpu["fast_math"] = fm
    return {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mcell-updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": cpu,
            "e2e": {"value": r["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


# ------------------------------------------------------------------------------------------------------------------
# verification of the timed run against the oracle (outside the timed region)
# ------------------------------------------------------------------------------------------------------------------
def verify(cfg, size, boxes, listeners, emitter_cells, outputs, gy):
    """outputs: [sources, emitters, 8] of the TIMED run.  The oracle solves every source; obstruction, wet gain, RT60 and both
    direction vectors must be bit-identical (up to the sign of zero), the low-pass cutoff within 1 ulp (2.5e-7)."""
    from oracle import pvoracle
    t0 = time.perf_counter()
    ora = pvoracle.OracleSim(size, size, cfg["resolution"], T=cfg["T"], efree=EFREE_275)
    for b in boxes:
        ora.add_aabb(*b)
    bad = []
    for i, l in enumerate(listeners):
        ora.results[:] = 0
        ora.generate(l)
        ora.analyze(l)
        for e, rc in enumerate(emitter_cells):
            got = outputs[i, e]
            if rc is None:
                ok = bool((got == -1.0).all())
            else:
                ref = ora.results[rc[0] * gy + rc[1]]
                exact = [0, 1, 2, 4, 5, 6, 7]
                same = (got[exact].view(np.uint32) == ref[exact].view(np.uint32)) | (got[exact] == ref[exact]) | (np.isnan(got[exact]) & np.isnan(ref[exact]))
                lp = abs(float(got[3]) - float(ref[3])) <= 2.5e-7 * max(abs(float(ref[3])), 1e-30)
                ok = bool(same.all() and lp)
            if not ok:
                bad.append((i, e))
    return {"ok": not bad, "sources": len(listeners), "emitters": len(emitter_cells), "mismatches": bad,
            "checker": "oracle/pv_oracle.c (pinned bit-for-bit to the unmodified reference, tests/test_oracle.py)",
            "seconds": time.perf_counter() - t0}


# ------------------------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs through the same library
# ------------------------------------------------------------------------------------------------------------------
def extra_latency(pvcuda, device):
    """Frame time (time steps + analyzer, CUDA events) of the reference's own contract case -- the Sandbox default 25 m world at
    resolution 275: 70 x 70 cells, 435 steps, one listener (PlaneverbSandbox/src/main.cpp:14-21) -- and of BASELINE configs[1]
    (Shoebox.pv, 512 x 512, 2000 steps)."""
    out = {}
    from planeverb_b200 import scenes as pscenes
    all_scenes = pscenes.load_scenes()
    for key, scene_name, n, T in (("contract_70", "FloorPlanScene", None, 0), ("config2", "Shoebox", 512, 2000)):
        if n is None:
            size, scale = 25.0, 1.0
        else:
            size, scale = pscenes.scaled_config(n)
        sc = pvcuda.Scene(size, size, 275, T=T, max_sources=1, device=device, efree=EFREE_275)
        for b in pscenes.boxes_of(all_scenes, scene_name, scale):
            sc.add_aabb(*b)
        listener = [(5.0 * scale, 0.0, 4.0 * scale)]
        best = None
        for it in range(6):
            sc.solve_async(listener)
            sc.wait()
            st, an, tot, nl = sc.timing()
            if it and (best is None or tot < best[2]):
                best = (st, an, tot, nl)
        out[key] = {"grid": [sc.gx, sc.gy], "time_steps": sc.T, "scene": scene_name, "frame_ms": best[2], "step_ms": best[0],
                    "analyzer_ms": best[1], "kernel_launches": best[3],
                    "Mcell_updates_per_s": sc.gx * sc.gy * sc.T / best[2] / 1e3}
        sc.close()
    return out


def extra_config4(pvcuda, sharding, device, rank, world, dist, tdev):
    """BASELINE configs[3]: HugeRoom.pv on 2048 x 2048, a FIXED list of 8 listener positions, 4000 steps, sharded over the N GPUs
    (strong scaling).  A 2048^2 source keeps 71 GB of pressure history, so a GPU solves its shard two sources at a time."""
    cfg = dict(scene="HugeRoom", n=2048, T=4000, sources=8, resolution=275)
    size, scale, boxes = scene_inputs(cfg)
    listeners = [((5.0 + 1.5 * i) * scale, 0.0, (4.0 + 0.75 * i) * scale) for i in range(cfg["sources"])]
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in EMITTERS]
    batch = 2
    plan = sharding.plan_batches(cfg["sources"], world, rank, batch)
    sc = pvcuda.Scene(size, size, 275, T=cfg["T"], max_sources=min(batch, max(1, max((hi - lo) for lo, hi in plan) if plan else 1)),
                      device=device, efree=EFREE_275)
    for b in boxes:
        sc.add_aabb(*b)
    sc.flush_geometry()

    def job():
        outs = []
        for lo, hi in plan:
            sc.solve_async(listeners[lo:hi])
            buf = pvcuda.pinned_array((hi - lo, len(emitters), 8))
            sc.lookup_wait(sc.lookup_async(emitters, buf, n=hi - lo))
            outs.append(np.array(buf, copy=True))
        return np.concatenate(outs) if outs else np.zeros((0, len(emitters), 8), np.float32)

    job()                                                  # warm
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = job()
    sc.wait()
    local_s = time.perf_counter() - t0
    allout = np.concatenate(sharding.gather_outputs(mine, dist, tdev, n_total=cfg["sources"]))
    secs = float(sharding.max_over_ranks([local_s], dist, tdev)[0])
    sc.close()
    units = cfg["n"] * cfg["n"] * cfg["T"] * cfg["sources"]
    return {"workload": workload_text(cfg, "in all") + f", sharded over {world} GPU(s), batches of <= {batch} sources", "scaling": "strong",
            "job_ms": secs * 1e3, "Mcell_updates_per_s": units / secs / 1e6,
            "outputs_checksum": float(np.nan_to_num(allout.astype(np.float64)).sum())}


def extra_config4_streamed(pvcuda, device, strong):
    """BASELINE configs[3] on ONE GPU as ONE batch: the streamed solver (pvc_create_streamed) keeps an 800-sample pressure
    history instead of 4000 (14 GB per source instead of 71), solves the response in 5 chunks and recomputes 4 of them for the
    backward Schroeder pass -- all eight sources fit the device at once.  Same job, listeners and emitters as config4_strong:
    the output checksums must be equal (the two paths are bit-identical, tests/test_gpu_streamed.py)."""
    cfg = dict(scene="HugeRoom", n=2048, T=4000, sources=8, resolution=275)
    history = 800
    need = pvcuda.memory_requirement(cfg["n"], cfg["n"], cfg["T"], cfg["sources"], history_steps=history)
    free, _ = pvcuda.device_memory(device)
    if need > 0.95 * free:
        return {"skipped": f"needs {need / 1e9:.0f} GB of device memory, {free / 1e9:.0f} GB free"}
    size, scale, boxes = scene_inputs(cfg)
    listeners = [((5.0 + 1.5 * i) * scale, 0.0, (4.0 + 0.75 * i) * scale) for i in range(cfg["sources"])]
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in EMITTERS]
    sc = pvcuda.Scene(size, size, 275, T=cfg["T"], max_sources=cfg["sources"], device=device, efree=EFREE_275, history_steps=history)
    for b in boxes:
        sc.add_aabb(*b)
    sc.flush_geometry()
    buf = pvcuda.pinned_array((cfg["sources"], len(emitters), 8))
    best = None
    for it in range(2):
        t0 = time.perf_counter()
        sc.solve_async(listeners)
        sc.lookup_wait(sc.lookup_async(emitters, buf, n=cfg["sources"]))
        sc.wait()
        secs = time.perf_counter() - t0
        if it:
            best = secs
    launches = sc.timing()[3]
    variant = sc.step_variant()
    sc.close()
    units = cfg["n"] * cfg["n"] * cfg["T"] * cfg["sources"]
    checksum = float(np.nan_to_num(np.array(buf, copy=True).astype(np.float64)).sum())
    return {"workload": workload_text(cfg, "in all") + f", one GPU, ONE batch of 8 on the streamed solver (history {history} samples, 5 chunks)",
            "job_ms": best * 1e3, "Mcell_updates_per_s": units / best / 1e6, "kernel_launches": launches, "step_kernel_variant": variant,
            "device_memory_GB": need / 1e9, "full_history_GB": pvcuda.memory_requirement(cfg["n"], cfg["n"], cfg["T"], cfg["sources"]) / 1e9,
            "outputs_checksum": checksum, "checksum_equals_config4_strong": bool(checksum == strong.get("outputs_checksum"))}


def extra_config5(pvcuda, sharding, device, rank, world, dist, tdev, frames=12):
    """BASELINE configs[4]: FloorPlanScene.pv on 1024 x 1024 with one AABB moved every frame (UpdateGeometry = Remove(old) +
    Add(new), re-voxelised on the device), 8 listener positions sharded over the N GPUs, contract response length (435 steps);
    per frame: edit + solve + emitter lookup + the all-gather of the per-emitter parameters."""
    cfg = dict(scene="FloorPlanScene", n=1024, T=435, sources=8, resolution=275)
    size, scale, boxes = scene_inputs(cfg)
    listeners = sharding.shard([((5.0 + 1.5 * i) * scale, 0.0, (4.0 + 0.75 * i) * scale) for i in range(cfg["sources"])], world, rank)
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in EMITTERS]
    S = max(1, len(listeners))
    sc = pvcuda.Scene(size, size, 275, T=cfg["T"], max_sources=S, device=device, efree=EFREE_275)
    for b in boxes:
        sc.add_aabb(*b)
    moving = list(boxes[-1])
    bufs = [pvcuda.pinned_array((S, len(emitters), 8)) for _ in range(2)]

    def frame(k):
        old = tuple(moving)
        moving[0] = boxes[-1][0] + 0.05 * scale * ((k % 7) - 3)
        sc.remove_aabb(*old)
        sc.add_aabb(*moving)
        if listeners:
            sc.solve_async(listeners)
            sc.lookup_wait(sc.lookup_async(emitters, bufs[k & 1], n=len(listeners)))
        return sharding.gather_outputs(bufs[k & 1][:len(listeners)], dist, tdev, n_total=cfg["sources"])

    for k in range(3):
        frame(k)
    if dist is not None:
        import torch
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(frames):
        out = frame(3 + k)
    sc.wait()
    local_s = time.perf_counter() - t0
    secs = float(sharding.max_over_ranks([local_s], dist, tdev)[0])
    sc.close()
    return {"workload": workload_text(cfg, "in all") + f", one AABB moved + re-voxelised per frame, sources sharded over {world} GPU(s)",
            "frames": frames, "frame_ms": secs * 1e3 / frames,
            "Mcell_updates_per_s": cfg["n"] * cfg["n"] * cfg["T"] * cfg["sources"] * frames / secs / 1e6,
            "outputs_checksum": float(np.nan_to_num(np.concatenate(out).astype(np.float64)).sum())}


# ------------------------------------------------------------------------------------------------------------------
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
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-row", action="store_true")
    ap.add_argument("--verify", dest="verify", action="store_true", default=None, help="check the timed run's outputs against the oracle (default at N = 1)")
    ap.add_argument("--no-verify", dest="verify", action="store_false")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_line(args)), flush=True)
        return 0

    scrubbed = sorted(k for k in os.environ if k.startswith("PVC_"))
    for k in scrubbed:
        del os.environ[k]                  # no tuning knob can reach the measured library (the release build reads none anyway)

    peak, peak_src = load_peaks()
    cfg = dict(WORKLOAD, n=args.n, T=args.T, sources=args.sources)
    config = {"workload": workload_text(cfg) + " (BASELINE.json configs[2])",
              "grid": [cfg["n"], cfg["n"]], "time_steps": cfg["T"], "sources_per_gpu": cfg["sources"],
              "l2_policy": "inputs larger than L2: every step streams the pressure history "
                           f"({4 * cfg['n'] * cfg['n'] * cfg['T'] * cfg['sources'] / 1e9:.0f} GB) through HBM",
              "parallelism": f"sources sharded, {world} rank(s), no data-path collective"}

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

    size, scale, boxes = scene_inputs(cfg)
    S = cfg["sources"]
    listeners = sharding.shard(bench_listeners(S * world, scale), world, rank)      # S sources on every rank
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in EMITTERS]
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
    local_outputs = np.array(em_bufs[(args.steps - 1) & 1], copy=True)      # this rank's emitter outputs of the last TIMED solve
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
    kernel_variant = scene.step_variant() if hasattr(scene, "step_variant") else args.variant
    gx, gy, T_run = scene.gx, scene.gy, scene.T
    scene.close()
    del bufs

    # ---------------- outside the timed region: verification and the other configs ----------------
    do_verify = args.verify if args.verify is not None else (world == 1)
    verdict = None
    if do_verify and rank == 0:
        verdict = verify(cfg, size, boxes, listeners, em_cells, local_outputs, gy)
    extras = {}
    if not args.no_extras and args.step_kernel == 0 and args.variant == 0:
        if rank == 0:
            extras.update(extra_latency(pvcuda, device))
        extras["config4_strong"] = extra_config4(pvcuda, sharding, device, rank, world, dist, tdev)
        if world == 1:
            extras["config4_streamed"] = extra_config4_streamed(pvcuda, device, extras["config4_strong"])
        extras["config5_dynamic"] = extra_config5(pvcuda, sharding, device, rank, world, dist, tdev)

    rc = 0
    if rank == 0:
        value = total_units / dev_s / 1e6
        step_gbs = ALGO_BYTES_PER_CELL_UPDATE * units_per_step * args.steps / step_s / 1e9
        step_launches = max(step_launches, 1)        # step-kernel launches of the timed solves (the rest of gpu_launches: analyzer kernels)
        cell_updates_per_launch = units_per_step * args.steps / step_launches
        avg_launch_s = step_s / step_launches
        kernel_name = {47: "pvc::ws2::stepKernel<14,4,1,false,true> (warp-specialised generational kernel, producer + publisher warps, up to 256 x 4 time steps per launch)",
                       50: "pvc::ws2::stepKernel<8,4,1,false,true> (generational kernel, 32-row tiles)",
                       18: "pvc::fusedStepKernel<8,6,2> (one launch per 4 time steps)",
                       60: "pvc::res::residentKernel<8,4,2>", 61: "pvc::res::residentKernel<10,4,2>", 62: "pvc::res::residentKernel<12,4,2>",
                       63: "pvc::res::residentKernel<16,4,1>", 64: "pvc::res::residentKernel<20,4,1>", 65: "pvc::res::residentKernel<18,4,1>", 66: "pvc::res::residentKernel<16,5,1>"}.get(kernel_variant, f"step kernel variant {kernel_variant}")
        if kernel_variant in (60, 61, 62, 63, 64, 65, 66):
            kernel_name += " (resident kernel: one launch per source batch runs all time steps with the tile state in registers; only halo strips pass through L2)"
        if args.step_kernel == 1:
            kernel_name = "pvc::baselinePressureKernel + pvc::baselineVelocityKernel"
        # physical roofs from the committed ncu capture of this kernel (profiles/step_kernel_ncu.json: DRAM bytes and issue-slot
        # utilisation per cell-update, scaled to the launches of THIS run); null when the capture is of another kernel
        ncu = None
        if os.path.exists(KERNEL_NCU) and args.step_kernel == 0:
            try:
                cand = json.load(open(KERNEL_NCU))
                if int(cand.get("variant", -1)) == int(kernel_variant) and int(cand.get("grid_n", cfg["n"])) == cfg["n"]:
                    ncu = cand
            except Exception:
                ncu = None
        traffic = ncu["dram_bytes_per_cell_update"] * cell_updates_per_launch if ncu else None
        # what the kernel must move at the very least: the 4-byte history record per cell-step (written once)
        traffic_floor = 4.0 * cell_updates_per_launch
        roofline = {"bound": "hbm", "kernel": kernel_name,
                    "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "physical_frac": (traffic / avg_launch_s / 1e9 / peak) if traffic else None,
                    "physical_floor_frac": traffic_floor / avg_launch_s / 1e9 / peak,
                    "issue_frac": (ncu["issue_active_pct"] / 100.0) if ncu else None,
                    "ncu_capture": (ncu.get("source") if ncu else None),
                    "launches_per_solve": step_launches / args.steps,
                    "algorithmic_bytes_per_launch": ALGO_BYTES_PER_CELL_UPDATE * cell_updates_per_launch,
                    "avg_launch_us": avg_launch_s * 1e6,
                    "note": "achieved = 28 B per cell-update x cell-updates of the timed steps / CUDA-event time of the step-kernel phase; above 1.0 "
                            "because the state never leaves the chip between time steps (temporal blocking / register residency): HBM is not this "
                            "kernel's roof.  physical_frac = DRAM bytes (ncu capture of the same kernel, per cell-update) / launch time / peak; "
                            "physical_floor_frac = the 4-byte history record alone; issue_frac = smsp__issue_active of that capture -- the roof that binds"}
        line = {
            "metric": METRIC,
            "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": total_units / e2e_s / 1e6, "unit": "Mcell-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "phases_ms_per_step": {"step_kernels": step_s * 1e3 / args.steps, "analyzer": ana_s * 1e3 / args.steps},
            "wall_ms_per_step": (t_wall1 - t_wall0) * 1e3 / args.steps,
            "outputs_checksum": float(np.nan_to_num(gathered.astype(np.float64)).sum()),
            "step_kernel_variant": kernel_variant,
            "env_scrubbed": scrubbed,
        }
        if verdict is not None:
            line["verified"] = bool(verdict["ok"])
            line["verify"] = verdict
            if not verdict["ok"]:
                rc = 1
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            r = run_reference_sample(1, 0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if not args.no_fast_row:
                fm = fast_math_row()
                if fm:
                    line["cpu_baseline"]["fast_math"] = fm
        print(json.dumps(line), flush=True)
        if rc:
            print("bench: VERIFICATION FAILED -- the timed run's outputs differ from the oracle: " + json.dumps(verdict["mismatches"]), file=sys.stderr, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
