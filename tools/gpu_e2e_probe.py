import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from planeverb_b200 import pvcuda
cfg = dict(bench.WORKLOAD)
size, scale, boxes, common = bench.scene_inputs(cfg)
S = 4
listeners = bench.bench_listeners(S, scale)
scene = pvcuda.Scene(size, size, 275, T=cfg['T'], max_sources=S)
cells = scene.gx * scene.gy
res = pvcuda.pinned_array((S, cells, 8)); dly = pvcuda.pinned_array((S, cells))
for it in range(8):
    t0 = time.perf_counter()
    scene.clear_geometry()
    for b in boxes: scene.add_aabb(*b)
    scene.flush_geometry()
    t1 = time.perf_counter()
    scene.solve_async(listeners); scene.wait()
    t2 = time.perf_counter()
    for i in range(S):
        pvcuda._check(pvcuda.lib().pvc_fetch_results(scene._solver, i, pvcuda._p(res[i]), pvcuda._p(dly[i])), 'fetch')
    t3 = time.perf_counter()
    out = [scene.lookup((5*scale, 0, 6*scale), s) for s in range(S)]
    t4 = time.perf_counter()
    print(f'frame {it}: geometry {1e3*(t1-t0):.2f} ms, solve {1e3*(t2-t1):.2f} ms, d2h {1e3*(t3-t2):.2f} ms, lookups {1e3*(t4-t3):.2f} ms, device {scene.timing()[:3]}')
