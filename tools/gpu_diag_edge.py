"""Diagnostic: the n=120 edge case (grid exactly one tile wide) on the resident variants, several repetitions, first bad sample."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
from planeverb_b200 import pvcuda
from oracle import pvoracle

def run(n, T, cell, variant, reps=3):
    size, _ = common.scaled_config(n)
    ora = pvoracle.OracleSim(size, size, 275, T=T, efree=0.0447895788)
    dx = float(ora.dx)
    L = ((cell[0] + 0.5) * dx, 0.0, (cell[1] + 0.5) * dx)
    boxes = [(-0.5 * dx, 0.3 * n * dx, 3 * dx, 4 * dx, 0.9), (n * dx, 0.7 * n * dx, 4 * dx, 6 * dx, 0.5), (0.6 * n * dx, n * dx, 5 * dx, 2.5 * dx, 0.97)]
    for b in boxes: ora.add_aabb(*b)
    ora.generate(L); ora.analyze(L)
    for rep in range(reps):
        gpu = pvcuda.Scene(size, size, 275, T=T, efree=0.0447895788, variant=variant)
        for b in boxes: gpu.add_aabb(*b)
        res, dly = gpu.solve([L])
        first = None
        for t in range(T):
            ok = common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(n + 1, n + 1))
            if not ok.all():
                bad = np.argwhere(~ok)
                first = (t, len(bad), bad[:6].tolist())
                break
        print(f"n={n} T={T} cell={cell} var={gpu.step_variant()} rep {rep}: " + ("all planes OK" if first is None else f"first bad t={first[0]} ({first[1]} cells) at {first[2]}"), flush=True)
        gpu.close()

for var in (0, 60, 61, 63):
    run(120, 97, (119, 119), var, 1)
run(120, 97, (60, 60), 60, 2)
run(240, 97, (239, 239), 60, 2)
run(121, 150, (60, 120), 60, 2)
run(360, 97, (359, 359), 61, 1)
run(360, 97, (359, 359), 66, 1)
run(320, 97, (319, 319), 61, 1)
