"""Small solve for compute-sanitizer / quick checks:  python tools/gpu_small_case.py VARIANT [SCENE|none] [N] [S] [T] [HISTORY]
N = 0: the Sandbox default 25 m world (70 x 70 cells); otherwise the scene scaled to N x N cells; HISTORY > 0: the streamed solver."""
import sys
sys.path.insert(0, '/root/repo')
from tests import common
from planeverb_b200 import pvcuda
scenes = common.load_scenes()
var = int(sys.argv[1]) if len(sys.argv) > 1 else 7
scene = sys.argv[2] if len(sys.argv) > 2 else 'SmallRoom'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
S = int(sys.argv[4]) if len(sys.argv) > 4 else 1
T = int(sys.argv[5]) if len(sys.argv) > 5 else 0
H = int(sys.argv[6]) if len(sys.argv) > 6 else 0
size, scale = (25.0, 1.0) if n == 0 else common.scaled_config(n)
G = pvcuda.Scene(size, size, 275, T=T, max_sources=S, variant=var, history_steps=H, efree=(0.0447895788 if H else -1.0))
if scene != 'none':
    for b in common.boxes_of(scenes, scene, scale): G.add_aabb(*b)
for _ in range(3):
    G.solve(common.listeners_for(S, scale), fetch=False)
print(var, scene, n, S, H, G.timing())
