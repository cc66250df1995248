"""Resident step kernel (variants 60..64): parity against the oracle on small cases, then step-phase timing against the
generational kernels.  python tools/gpu_res_check.py [parity] [time]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
from planeverb_b200 import pvcuda
from oracle import pvoracle

scenes = common.load_scenes()
what = sys.argv[1:] or ["parity", "time"]


def parity(scene, n, T, variant, S=1, res=275, listeners=None):
    if n is None:
        size, scale = 25.0, 1.0
    else:
        size, scale = common.scaled_config(n, res)
    Ls = listeners or common.listeners_for(S, scale)
    try:
        G = pvcuda.Scene(size, size, res, T=T, max_sources=len(Ls), variant=variant)
    except pvcuda.PlaneverbCudaError as e:
        print(f"  {scene} n={n} T={T} var={variant}: create failed: {e}")
        return False
    ok_all = True
    boxes = common.boxes_of(scenes, scene, scale) if scene else []
    for b in boxes:
        G.add_aabb(*b)
    try:
        res_, dly = G.solve(Ls)
    except pvcuda.PlaneverbCudaError as e:
        print(f"  {scene} n={n} T={T} var={variant}: solve failed: {e}")
        G.close()
        return False
    for i, l in enumerate(Ls):
        O = pvoracle.OracleSim(size, size, res, T=T)
        for b in boxes:
            O.add_aabb(*b)
        O.generate(l)
        O.analyze(l)
        bad_planes = [t for t in sorted(set([0, 1, 3, 4, 7, 8, G.T // 2, G.T - 2, G.T - 1]))
                      if not common.bit_equal(G.pressure(t, source=i), O.hist[t].reshape(G.gx + 1, G.gy + 1)).all()]
        p, vx, vy = G.state(source=i)
        shp = (G.gx + 1, G.gy + 1)
        st_ok = (common.bit_equal(p, O.p.reshape(shp)).all(), common.bit_equal(vx, O.vx.reshape(shp)).all(), common.bit_equal(vy, O.vy.reshape(shp)).all())
        dl_ok = np.array_equal(dly[i], O.delay)
        valid = (O.delay < 3e38) & ~O.clamped.astype(bool)
        fields_ok = [bool((common.bit_equal(res_[i][valid, k], O.results[valid, k]) | (np.isnan(res_[i][valid, k]) & np.isnan(O.results[valid, k]))).all())
                     for k in (0, 1, 2, 4, 5, 6, 7)]
        ok = (not bad_planes) and all(st_ok) and dl_ok and all(fields_ok)
        ok_all &= ok
        print(f"  {scene} n={G.gx} T={G.T} var={variant} src {i}: {'OK' if ok else 'MISMATCH'} planes_bad={bad_planes} state={st_ok} delay={dl_ok} fields={fields_ok}", flush=True)
    G.close()
    return ok_all


def timing(scene, n, T, S, variants, res=275):
    if n is None:
        size, scale = 25.0, 1.0
    else:
        size, scale = common.scaled_config(n, res)
    Ls = common.listeners_for(S, scale)
    for var in variants:
        try:
            G = pvcuda.Scene(size, size, res, T=T, max_sources=S, variant=var, efree=0.0447895788)
        except pvcuda.PlaneverbCudaError as e:
            print(f"{scene} {n} x{S} T={T} var={var}: create failed: {e}")
            continue
        for b in (common.boxes_of(scenes, scene, scale) if scene else []):
            G.add_aabb(*b)
        best = None
        try:
            for it in range(5):
                t0 = time.perf_counter()
                G.solve(Ls, fetch=False)
                wall = (time.perf_counter() - t0) * 1e3
                st, an, tot, nl = G.timing()
                if it and (best is None or st < best[0]):
                    best = (st, an, tot, nl, wall)
        except pvcuda.PlaneverbCudaError as e:
            print(f"{scene} {n} x{S} T={T} var={var}: solve failed: {e}")
            G.close()
            continue
        st, an, tot, nl, wall = best
        cu = G.gx * G.gy * G.T * S
        print(f"{scene} {G.gx}^2 x{S} T={G.T} var={var}: steps {st:.3f} ms ({cu / st / 1e6:.1f} Gcell/s) analyzer {an:.3f} ms total {tot:.3f} ms wall {wall:.3f} ms launches {nl}", flush=True)
        G.close()


if "parity" in what:
    print("== parity")
    ok = True
    for var in (60, 61, 62, 63, 64, 65, 66):
        ok &= parity("FloorPlanScene", 250, 301, var)
        ok &= parity("SmallRoom", None, 0, var)
    ok &= parity("BigRoom", 300, 122, 61, S=3)
    ok &= parity("Shoebox", 512, 200, 60)
    ok &= parity("HugeRoom", 384, 203, 62, S=2)
    ok &= parity(None, 257, 130, 63)
    ok &= parity("FloorPlanScene", 240, 97, 60, listeners=[(239.5 * 0.3565818, 0, 239.5 * 0.3565818), (100 * 0.3565818, 0, 239.5 * 0.3565818)])     # padding row in the bottom halo (240 = 10 x 24)
    ok &= parity("HugeRoom", 720, 203, 66, S=2)                                                                 # 720 = 10 x 72: the same with 5-row warps
    # listeners on a wall cell, on the padding row / column, in the corner cell
    size, scale = common.scaled_config(250)
    ok &= parity("FloorPlanScene", 250, 90, 61, listeners=[(1.05 * scale, 0, 4 * scale), (250.2 * 0.3565818 , 0, 4 * scale), (0.1, 0, 0.1), (4 * scale, 0, 250.3 * 0.3565818)])
    print("PARITY", "ALL OK" if ok else "FAILED")

if "time" in what:
    print("== timing")
    timing("FloorPlanScene", None, 0, 1, [50, 60, 61, 62, 63, 64])
    timing("Shoebox", 512, 2000, 1, [50, 60, 61, 62, 63, 64])
    timing("BigRoom", 1024, 1000, 1, [50, 61, 62, 64, 65])
    timing("BigRoom", 1024, 1000, 4, [47, 61, 62, 64, 65])
    timing("FloorPlanScene", 1024, 1000, 4, [47, 61, 64, 65])
    timing("HugeRoom", 768, 1000, 4, [47, 60, 61, 62, 63, 64])
    timing("HugeRoom", 256, 1000, 8, [47, 50, 60, 61, 62, 63, 64])
