"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/*.h declares, the host
logic (index/scalar derivation) matches the reference, and the product fails loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pvoracle
from planeverb_b200 import pvcuda
from tests import common

INCLUDE = os.path.join(common.ROOT, "include")


def _declared(header):
    text = open(os.path.join(INCLUDE, header)).read()
    return sorted(set(re.findall(r"PVC_API[^;(]*?\b(pv[cx]_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = pvcuda.lib()
    cuda_syms = _declared("planeverb_cuda.h")
    ext_syms = _declared("planeverb_ext.h")
    assert len(cuda_syms) >= 20 and len(ext_syms) >= 13
    for name in cuda_syms + ext_syms:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert sorted(pvcuda.PVC_SYMBOLS) == cuda_syms
    assert sorted(pvcuda.PVX_SYMBOLS) == ext_syms


def test_planeverb_c_abi_symbols_exported():
    """The Unity C ABI of PlaneverbUnity.cpp:12-135 and nothing renamed."""
    L = pvcuda.lib()
    for name in ["UnityPluginLoad", "UnityPluginUnload", "PlaneverbInit", "PlaneverbExit", "PlaneverbEmit",
                 "PlaneverbUpdateEmission", "PlaneverbEndEmission", "PlaneverbGetOutput", "PlaneverbAddGeometry",
                 "PlaneverbUpdateGeometry", "PlaneverbRemoveGeometry", "PlaneverbSetListenerPosition"]:
        assert hasattr(L, name), name


def test_no_device_fails_loudly():
    if pvcuda.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pvcuda.PlaneverbCudaError) as e:
        pvcuda.Scene(25.0, 25.0, 275)
    assert "no CUDA device" in str(e.value)


def test_invalid_config_rejected_like_the_reference():
    """PvContext.cpp:101-107: resolution < 275, zero size -> pv_InvalidConfig (here PVC_ERR_INVALID)."""
    for args in [(25.0, 25.0, 100), (0.0, 25.0, 275), (25.0, 0.0, 275)]:
        with pytest.raises(pvcuda.PlaneverbCudaError) as e:
            pvcuda.Scene(*args)
        assert "invalid" in str(e.value)


@pytest.mark.parametrize("res", [275, 375, 500, 750, 300])
def test_host_derivation_matches_oracle(res):
    for size in (25.0, 10.0, 37.3):
        cfg, dt, free_r, free = pvcuda.derive(res, size, size)
        dx, odt, fs = pvoracle.grid_params(res)
        gx, gy, T, courant = pvoracle.derived(res, size, size)
        assert (cfg.gx, cfg.gy, cfg.T, cfg.fs) == (gx, gy, T, fs)
        assert (np.float32(cfg.dx), dt, np.float32(cfg.courant)) == (dx, odt, courant)
        assert (cfg.flux_samples, cfg.dry_samples, cfg.wet_samples, cfg.tail_samples) == pvoracle.windows(fs)
        assert np.array_equal(pvcuda.derive_pulse(res, fs, T), pvoracle.gaussian_pulse(res, fs, T))
    cfg, _, _, _ = pvcuda.derive(res, 25.0, 25.0, T=500)
    assert cfg.T == 500


def test_host_rect_and_listener_cells_match_oracle(scenes):
    res = 275
    sim = pvoracle.OracleSim(25.0, 25.0, res)
    for name in ("FloorPlanScene", "HugeRoom", "DirectionTester"):
        for b in common.boxes_of(scenes, name):
            q = pvcuda.derive_rect(res, *b, add=True)
            before = sim.b.copy()
            sim.b[:] = 1
            sim.add_aabb(*b)
            grid = sim.b.reshape(sim.gx + 1, sim.gy + 1) == 0
            want = np.zeros_like(grid)
            want[max(q.r0, 0):max(min(q.r1, sim.gx + 1), 0), max(q.c0, 0):max(min(q.c1, sim.gy + 1), 0)] = True
            assert np.array_equal(grid, want), (name, b)
            assert np.float32(q.admittance) == (np.float32(1) - np.float32(b[4])) / (np.float32(1) + np.float32(b[4]))
            sim.b[:] = before
    for pos in [(5.0, 4.0), (0.0, 0.0), (24.9, 24.9), (12.345, 6.789), (0.3565, 0.3566)]:
        l = pvcuda.derive_listener(res, *pos)
        assert (l.cell_r, l.cell_c) == sim.listener_cell((pos[0], 0.0, pos[1]))
        inv = np.float32(1.0) / sim.dx
        assert (l.efree_r, l.efree_c) == (int(np.float32(pos[0]) * inv), int(np.float32(pos[1]) * inv))


def test_emitter_cell_bounds():
    """Analyzer.cpp:110-113; pos == gridSize is rejected here (the reference indexes outside the lattice)."""
    assert pvcuda.derive_emitter_cell(275, 25.0, 25.0, 5.0, 6.0) == (14, 16)
    assert pvcuda.derive_emitter_cell(275, 25.0, 25.0, 30.0, 6.0) is None
    assert pvcuda.derive_emitter_cell(275, 25.0, 25.0, -1.0, 6.0) is None
    dx = float(pvoracle.grid_params(275)[0])
    assert pvcuda.derive_emitter_cell(275, 25.0, 25.0, 70 * dx + 0.01, 1.0) is None
    assert pvcuda.derive_emitter_cell(275, 25.0, 25.0, 69.5 * dx, 1.0) == (69, 2)


def test_struct_layouts_match_header():
    assert C.sizeof(pvcuda.PvcConfig) == 15 * 4
    assert C.sizeof(pvcuda.PvcRect) == 24
    assert C.sizeof(pvcuda.PvcListener) == 24
