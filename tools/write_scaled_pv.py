"""Write a .pv scene (PlaneverbSandbox format: count, then `id posX posY width height absorption`) scaled from the 25 m
authoring world to an n x n cell grid at resolution 275:  python tools/write_scaled_pv.py FloorPlanScene 1024 out.pv"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import common
name, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
size, scale = common.scaled_config(n)
boxes = common.boxes_of(common.load_scenes(), name, scale)
with open(out, "w") as f:
    f.write(f"{len(boxes)}\n")
    for i, b in enumerate(boxes):
        f.write(f"{i} {b[0]:.9g} {b[1]:.9g} {b[2]:.9g} {b[3]:.9g} {b[4]:.9g}\n")
print(f"{size:.9g} {scale:.9g}")
