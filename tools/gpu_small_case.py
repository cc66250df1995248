import sys
sys.path.insert(0, '/root/repo')
from tests import common
from planeverb_b200 import pvcuda
scenes = common.load_scenes()
var = int(sys.argv[1]) if len(sys.argv) > 1 else 7
scene = sys.argv[2] if len(sys.argv) > 2 else 'SmallRoom'
G = pvcuda.Scene(25.0, 25.0, 275, max_sources=1, variant=var)
if scene != 'none':
    for b in common.boxes_of(scenes, scene): G.add_aabb(*b)
for _ in range(3):
    G.solve([common.DEFAULT_LISTENER], fetch=False)
print(var, scene, G.timing())
