"""Step-kernel timing per variant at BASELINE config 3 (1024^2, 4 sources): python tools/gpu_variant_bench.py 36 39 40 [T]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
from planeverb_b200 import pvcuda
scenes = common.load_scenes()
args = [int(a) for a in sys.argv[1:]]
variants = [a for a in args if a < 100] or [36]
T = ([a for a in args if a >= 100] or [4000])[0]
n = int(os.environ.get("PVB_N", "1024")); S = int(os.environ.get("PVB_S", "4")); scene = os.environ.get("PVB_SCENE", "BigRoom")
size, scale = common.scaled_config(n)
for var in variants:
    G = pvcuda.Scene(size, size, 275, T=T, max_sources=S, variant=var, efree=0.0447895788)
    for b in common.boxes_of(scenes, scene, scale): G.add_aabb(*b)
    Ls = common.listeners_for(S, scale)
    best = None
    for it in range(4):
        G.solve(Ls, fetch=False)
        st, an, tot, nl = G.timing()
        if it and (best is None or st < best[0]): best = (st, an, tot, nl)
    st, an, tot, nl = best
    cu = n * n * T * S
    print(f"{scene} {n}^2 x{S} T={T} var={var} order={os.environ.get('PVC_TILE_ORDER','cost')}: steps {st:.2f} ms ({cu/st/1e6:.1f} Gcell/s) analyzer {an:.2f} ms total {tot:.2f} ms launches {nl}", flush=True)
    G.close()
