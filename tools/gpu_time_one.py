"""Step-phase timing of one configuration: python tools/gpu_time_one.py SCENE N T S VARIANT [reps [history_steps]]
(N = 0: the 25 m contract grid; history_steps > 0: the streamed solver, -1: automatic)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from planeverb_b200 import pvcuda, scenes as pscenes

scene, n, T, S, var = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 4
history = int(sys.argv[7]) if len(sys.argv) > 7 else 0
size, scale = (25.0, 1.0) if n == 0 else pscenes.scaled_config(n)
G = pvcuda.Scene(size, size, 275, T=T, max_sources=S, variant=var, efree=0.0447895788, history_steps=history)
if scene != "none":
    for b in pscenes.boxes_of(pscenes.load_scenes(), scene, scale):
        G.add_aabb(*b)
Ls = [((5.0 + 1.5 * i) * scale, 0.0, (4.0 + 0.75 * i) * scale) for i in range(S)]
best = None
for it in range(reps):
    G.solve(Ls, fetch=False)
    st, an, tot, nl = G.timing()
    if it and (best is None or st < best[0]):
        best = (st, an, tot, nl)
st, an, tot, nl = best
cu = G.gx * G.gy * G.T * S
hs = f" history={G.history_steps}" if G.history_steps else ""
print(f"{scene} {G.gx}^2 x{S} T={G.T} var={G.step_variant()}{hs} dbg={os.environ.get('PVC_RES_DEBUG', '-')}: steps {st:.3f} ms ({cu / st / 1e6:.1f} Gcell/s) "
      f"analyzer {an:.3f} ms total {tot:.3f} ms launches {nl}", flush=True)
G.close()
