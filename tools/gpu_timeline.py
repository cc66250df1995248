import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
from tests import common
from planeverb_b200 import pvcuda
scenes = common.load_scenes()
size, scale = common.scaled_config(1024)
L = pvcuda.lib()
L.pvc_debug_timeline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
cases = [(1, 7), (4, 7), (4, 6), (4, 10)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for S, var in cases:
    G = pvcuda.Scene(size, size, 275, T=400, max_sources=S, variant=var, efree=0.0447895788)
    for b in common.boxes_of(scenes, 'BigRoom', scale): G.add_aabb(*b)
    Ls = common.listeners_for(S, scale)
    G.solve(Ls, fetch=False)
    buf = np.zeros((4096, 8), np.uint64)
    n = L.pvc_debug_timeline(G._solver, S, buf.ctypes.data_as(C.c_void_p), 4096)
    t = buf[:n].astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1000.0
    d = np.diff(rel[:, :7], axis=1)
    print(f'S={S} var={var} blocks={n}: kernel span {rel[:,6].max():.1f} us; per-CTA phases mean us: load {d[:,0].mean():.2f} steps {d[:,1:5].mean(axis=0).round(2)} store {d[:,5].mean():.2f} total {(rel[:,6]-rel[:,0]).mean():.2f}')
    starts = np.sort(rel[:, 0])
    print('   CTA start quantiles (us)', np.quantile(starts, [0, .25, .5, .75, 1]).round(1), ' end quantiles', np.quantile(rel[:, 6], [0, .25, .5, .75, 1]).round(1))
    print('   load-phase quantiles (us)', np.quantile(d[:, 0], [0, .1, .25, .5, .75, .9, 1]).round(2), ' first-step quantiles', np.quantile(d[:, 1], [0, .25, .5, .75, 1]).round(2))
    tot = rel[:, 6] - rel[:, 0]
    print('   CTA duration quantiles', np.quantile(tot, [0, .1, .5, .9, 1]).round(2))
    G.close()
