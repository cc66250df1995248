"""BASELINE.json configs[4]: FloorPlanScene.pv with one AABB moved every frame (UpdateGeometry = Remove + Add, re-voxelised on
the device) followed by a full 1024 x 1024 solve.  Per-frame latency through the host-buffer C-ABI (pvx_*), one GPU:
  python tools/gpu_config5.py [S] [T]      (S sources on this GPU; T = 435 is the contract's response length at res 275)
Prints, per mode, the steady-state frame time: geometry edit + solve + result copy, synchronous and pipelined one frame deep."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import common
from planeverb_b200 import pvcuda

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 435
n = 1024
scenes = common.load_scenes()
size, scale = common.scaled_config(n)
boxes = common.boxes_of(scenes, "FloorPlanScene", scale)
door = boxes[3]
G = pvcuda.Scene(size, size, 275, T=T, max_sources=S, efree=0.0447895788)
for b in boxes:
    G.add_aabb(*b)
cells = n * n
Ls = common.listeners_for(S, scale)
emitters = [(x * scale, 0.0, z * scale) for (x, z) in common.EMITTERS]
bufs = [(pvcuda.pinned_array((S, cells, 8)), pvcuda.pinned_array((S, cells))) for _ in range(2)]
eb = [pvcuda.pinned_array((S, len(emitters), 8)) for _ in range(2)]
cur = [door]

def move(k):
    moved = (door[0] + 0.05 * ((k % 16) + 1) * scale, door[1] - 0.03 * ((k % 16) + 1) * scale, door[2], door[3], door[4])
    G.remove_aabb(*cur[0]); G.add_aabb(*moved)
    cur[0] = moved

def run(mode, frames=12, warm=3):
    ts = []
    for k in range(frames + warm):
        if k == warm:
            G.wait(); t0 = time.perf_counter()
        move(k)
        if mode == "grids":                 # full result grids to the host every frame, synchronous
            G.solve(Ls, out=bufs[0])
        elif mode == "grids-pipelined":     # grids of frame k land while frame k+1 is solved
            G.solve_pipelined(Ls, bufs[k & 1])
        elif mode == "emitters":            # only the emitters' outputs leave the device (what GetOutput needs), synchronous
            G.solve_async(Ls); G.lookup_wait(G.lookup_async(emitters, eb[0]))
        elif mode == "emitters-pipelined":
            G.solve_async(Ls); t = G.lookup_async(emitters, eb[k & 1])
            if k: G.lookup_wait(prev)
            prev = t
    if mode == "grids-pipelined": G.fetch_wait()
    if mode == "emitters-pipelined": G.lookup_wait(prev)
    G.wait()
    dt = (time.perf_counter() - t0) / frames
    st, an, tot, nl = G.timing()
    print(f"config5 FloorPlanScene {n}^2 S={S} T={T} mode={mode:20s}: {1e3 * dt:8.3f} ms/frame ({1 / dt:7.1f} frames/s)  "
          f"device: steps {st:.3f} analyzer {an:.3f} ms, {nl} launches; {n * n * T * S / dt / 1e6:.0f} Mcell-updates/s", flush=True)

for mode in ("grids", "grids-pipelined", "emitters", "emitters-pipelined"):
    run(mode)
G.close()
