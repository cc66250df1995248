"""The reference's .pv scenes (imported to scenes.json by tools/import_scenes.py) and the BASELINE-style scaling of a
25 m authoring world onto an n x n grid.  Host arithmetic through the C-ABI's own derivation
(pvx_derive: the reference's single-precision expressions); needs no GPU and nothing from oracle/."""
import json
import os

import numpy as np

SCENES_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes.json")


def load_scenes():
    return json.load(open(SCENES_JSON))


def boxes_of(scenes, name, scale=1.0):
    """AABBs (posX, posY, width, height, absorption) of a .pv scene, optionally scaled from the 25 m authoring world."""
    out = []
    for b in scenes[name]["boxes"]:
        out.append((np.float32(b["pos"][0] * scale), np.float32(b["pos"][1] * scale),
                    np.float32(b["width"] * scale), np.float32(b["height"] * scale), np.float32(b["absorption"])))
    return [tuple(float(v) for v in t) for t in out]


def size_for_cells(resolution, n):
    """gridSizeInMeters that Grid's constructor truncates to exactly n x n cells (Grid.cpp:48-49): (n + 0.5) * dx."""
    from planeverb_b200 import pvcuda
    cfg, _, _, _ = pvcuda.derive(resolution, 1.0, 1.0)
    size = float(np.float32((n + 0.5) * float(np.float32(cfg.dx))))
    chk, _, _, _ = pvcuda.derive(resolution, size, size)
    if (chk.gx, chk.gy) != (n, n):
        raise ValueError(f"size {size} truncates to {chk.gx} x {chk.gy} cells, wanted {n}")
    return size


def scaled_config(n, resolution=275):
    """(size_m, scale): the grid size that yields exactly n x n cells and the factor that makes a 25 m scene fill it."""
    from planeverb_b200 import pvcuda
    cfg, _, _, _ = pvcuda.derive(resolution, 1.0, 1.0)
    return size_for_cells(resolution, n), n * float(np.float32(cfg.dx)) / 25.0
