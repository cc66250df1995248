import sys, time, json, numpy as np
sys.path.insert(0, '/root/repo')
from tests import common
from planeverb_b200 import pvcuda
scenes = common.load_scenes()
size, scale = common.scaled_config(1024)
for S in (1, 2, 4):
    for var in (18, 22, 30, 33, 36, 37, 38):
        G = pvcuda.Scene(size, size, 275, T=1000, max_sources=S, variant=var, efree=0.0447895788)
        for b in common.boxes_of(scenes, 'BigRoom', scale): G.add_aabb(*b)
        Ls = common.listeners_for(S, scale)
        for _ in range(4): G.solve(Ls, fetch=False)
        st, an, tot, nl = G.timing()
        cu = 1024*1024*1000*S
        print(f'S={S} var={var}: steps {st:.2f} ms ({cu/st/1e6:.1f} Gcell/s) analyzer {an:.2f} ms', flush=True)
        G.close()
