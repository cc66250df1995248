"""Import the reference's .pv scene fixtures (text: count, then `id posX posY width height absorption`
per line -- PlaneverbSandbox/src/Editor/Editor.cpp:245-281) into planeverb_b200/scenes/scenes.json.

Run in the build container only (reads /root/reference); the JSON travels with the repo so neither
tests nor bench.py need the reference tree at run time."""
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
FILES = ["SmallRoom.pv", "Shoebox.pv", "BigRoom.pv", "HugeRoom.pv", "SingleWall.pv", "DirectionTester.pv",
         "ExampleProject.pv", "DemoFiles/FloorPlanScene.pv", "DemoFiles/MiddleWallScene.pv",
         "DemoFiles/SmallRoomScene.pv", "DemoFiles/UnityReplicationTest.pv"]
out = {}
for f in FILES:
    toks = open(os.path.join(REF, f)).read().split()
    n = int(toks[0])
    boxes = []
    for i in range(n):
        t = toks[1 + 6 * i: 7 + 6 * i]
        boxes.append({"id": int(t[0]), "pos": [float(t[1]), float(t[2])], "width": float(t[3]),
                      "height": float(t[4]), "absorption": float(t[5])})
    out[os.path.splitext(os.path.basename(f))[0]] = {"source": f, "world_m": 25.0, "boxes": boxes}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "planeverb_b200", "scenes", "scenes.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, {k: len(v["boxes"]) for k, v in out.items()})
