"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled in place (oracle/_ref/libpvref.so,
built by oracle/refdriver/Makefile from /root/reference).  Run in the build container only; the
fixtures travel with the repo so neither the CPU nor the GPU tests need the reference tree.

Each fixture holds, for one (scene, grid, resolution, T, listener) case: the inputs, the reference's
coefficient fields, a few pressure/velocity snapshots, the impulse response of one probe cell, the
onset delays and all eight analyzer outputs of every cell.  The reference ships no golden vectors of
its own; these are outputs of the reference itself.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pvref            # noqa: E402
from tests import common            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name, scene, n (None = native 25 m world), resolution, T override, listener (pre-scale metres)
CASES = [
    ("smallroom_70", "SmallRoom", None, 275, 0, (5, 0, 4)),
    ("bigroom_70", "BigRoom", None, 275, 0, (5, 0, 4)),
    ("shoebox_70", "Shoebox", None, 275, 0, (5, 0, 4)),
    ("hugeroom_70", "HugeRoom", None, 275, 0, (5, 0, 4)),
    ("floorplan_70", "FloorPlanScene", None, 275, 0, (5, 0, 4)),
    ("singlewall_95_res375", "SingleWall", None, 375, 0, (12.5, 0, 12.5)),
    ("middlewall_127_res500", "MiddleWallScene", None, 500, 0, (6.5, 0, 4.25)),
    ("smallroom_128_T500", "SmallRoom", 128, 275, 500, (5, 0, 4)),          # BASELINE.json configs[0]
    ("directiontester_101_T300", "DirectionTester", 101, 275, 300, (20.3, 0, 3.1)),
]


def run_case(name, scene, n, res, T, listener, scenes):
    if n is None:
        size, scale = 25.0, 1.0
    else:
        size, scale = common.scaled_config(n, res)
    L = tuple(float(np.float32(v * scale)) for v in listener)
    boxes = common.boxes_of(scenes, scene, scale)
    sim = pvref.RefSim(size, size, res, T=T)
    assert sim.pulse_mismatch == 0
    for b in boxes:
        sim.add_aabb(*b)
    b, R = sim.coef()
    sim.generate(L)
    sim.analyze(L)
    results, delay = sim.results()
    snaps_t = sorted(set([0, 1, 2, 3, 4, 5, 8, 20, 57, sim.T // 2, sim.T - 1]))
    snaps = [sim.snapshot(t) for t in snaps_t]
    pr, pc = sim.gx // 2, sim.gy // 3
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        meta=json.dumps(dict(scene=scene, size=size, scale=scale, resolution=res, T=sim.T, T_override=T,
                             gx=sim.gx, gy=sim.gy, fs=sim.fs, listener=L, probe=[pr, pc], snaps_t=snaps_t)),
        boxes=np.array(boxes, np.float32), b=b.astype(np.int8), R=R, pulse=sim.pulse(),
        scalars=np.array([sim.dx, sim.dt, sim.efree, sim.courant], np.float32),
        snap_p=np.stack([s[0] for s in snaps]), snap_vx=np.stack([s[1] for s in snaps]),
        snap_vy=np.stack([s[2] for s in snaps]),
        ir=sim.ir(pr, pc), results=results, delay=delay)
    valid = delay < 3e38
    print(f"{name}: {sim.gx}x{sim.gy} T={sim.T} valid={int(valid.sum())} sum_occ={results[valid, 0].astype(np.float64).sum():.6f}")
    sim.close()


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    scenes = common.load_scenes()
    for case in CASES:
        run_case(*case, scenes)
