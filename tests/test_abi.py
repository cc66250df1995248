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


@pytest.mark.parametrize("gen_chunk,src_group,num_gen,nsrc,tps", [
    (16, 2, 40, 4, 6),        # the bench configuration's shape (two groups of two sources), a partial last chunk
    (16, 3, 16, 3, 5),        # one group, exactly one chunk
    (4, 2, 9, 5, 4),          # uneven last group (2 + 2 + 1) and a partial last chunk
    (1, 1, 5, 3, 7),          # degenerate: generation-major over single-source groups
    (16, 1, 3, 1, 11),        # single source, fewer generations than a chunk
])
def test_ws2_work_item_order_is_a_dependency_respecting_bijection(gen_chunk, src_group, num_gen, nsrc, tps):
    """The generational step kernel pulls work items from one counter and waits for an item's dependencies: the tile and
    its neighbours, same source, previous generation.  That loop cannot deadlock if (a) every (source, generation, tile) is
    exactly one item and (b) every item of generation g-1 of a source precedes every item of generation g of that source.
    pvc_debug_ws2_item evaluates the very inline function the kernel calls (pvc_internal.h::ws2DecodeItem) on the host."""
    import ctypes as C
    L = pvcuda.lib()
    out = (C.c_int * 3)()
    total = num_gen * nsrc * tps
    seen = {}
    last_of = {}            # (source, generation) -> largest item index
    first_of = {}           # (source, generation) -> smallest item index
    for w in range(total):
        assert L.pvc_debug_ws2_item(w, gen_chunk, src_group, num_gen, nsrc, tps, out) == 0
        s, g, o = out[0], out[1], out[2]
        assert 0 <= s < nsrc and 0 <= g < num_gen and 0 <= o < tps
        assert (s, g, o) not in seen
        seen[(s, g, o)] = w
        last_of[(s, g)] = w
        first_of.setdefault((s, g), w)
    assert len(seen) == total
    for s in range(nsrc):
        for g in range(1, num_gen):
            assert last_of[(s, g - 1)] < first_of[(s, g)]
    # sources of one group interleave inside a generation (their state shares the L2), groups do not inside a chunk
    for w in (-1, total):
        assert L.pvc_debug_ws2_item(w, gen_chunk, src_group, num_gen, nsrc, tps, out) == pvcuda.PVC_ERR_INVALID
    assert L.pvc_debug_ws2_item(0, gen_chunk, nsrc + 1, num_gen, nsrc, tps, out) == pvcuda.PVC_ERR_INVALID


@pytest.mark.parametrize("gen_chunk,src_group,num_gen,nsrc,tx,ty,band", [
    (8, 1, 20, 2, 3, 11, 4),       # config 4's shape in small: single-source groups, three bands, a partial last chunk
    (8, 1, 8, 1, 2, 7, 7),         # one band = the whole grid: plain generation-major order
    (4, 2, 9, 3, 2, 9, 2),         # bands shorter than a chunk (a band is empty in late generations), uneven last group
    (16, 1, 5, 1, 4, 5, 3),        # fewer generations than a chunk
])
def test_ws2_banded_item_order_respects_the_tile_dependencies(gen_chunk, src_group, num_gen, nsrc, tx, ty, band):
    """The banded order of the generational kernel (pvc_internal.h::Ws2Order: bands of tile rows that shift up one row per
    generation and run a whole chunk of generations before the next band starts) gives up "generation g-1 entirely before
    generation g" -- what must still hold is the kernel's actual wait condition: every (source, generation, tile) is exactly
    one item, and the tile and its up-to-8 neighbours of the previous generation precede it."""
    import ctypes as C
    L = pvcuda.lib()
    out = (C.c_int * 3)()
    total = num_gen * nsrc * tx * ty
    at = {}
    for w in range(total):
        assert L.pvc_debug_ws2_item_banded(w, gen_chunk, src_group, num_gen, nsrc, tx, ty, band, out) == 0
        key = (out[0], out[1], out[2])
        assert 0 <= key[0] < nsrc and 0 <= key[1] < num_gen and 0 <= key[2] < tx * ty and key not in at
        at[key] = w
    assert len(at) == total
    for (s, g, o), w in at.items():
        if g == 0:
            continue
        r, c = divmod(o, tx)
        for dr in (-1, 0, 1):
            for dc in (-1, 0, 1):
                rr, cc = r + dr, c + dc
                if 0 <= rr < ty and 0 <= cc < tx:
                    assert at[(s, g - 1, rr * tx + cc)] < w, (s, g, r, c, rr, cc)
    # a band really runs several generations before the next band starts: the order is not generation-major
    if band < ty and num_gen > 1:
        first_gen1 = min(w for (s, g, o), w in at.items() if g == 1 and s == 0)
        last_gen0 = max(w for (s, g, o), w in at.items() if g == 0 and s == 0)
        assert first_gen1 < last_gen0
    assert L.pvc_debug_ws2_item_banded(total, gen_chunk, src_group, num_gen, nsrc, tx, ty, band, out) == pvcuda.PVC_ERR_INVALID
    assert L.pvc_debug_ws2_item_banded(0, gen_chunk, src_group, num_gen, nsrc, tx, ty, 0, out) == pvcuda.PVC_ERR_INVALID


def test_product_scene_scaling_matches_the_oracle():
    """bench.py sizes its scenes with the product's own helper (planeverb_b200.scenes, through pvx_derive); it must agree with
    the oracle's derivation to the bit for every BASELINE grid size."""
    from planeverb_b200 import scenes as pscenes
    for n in (70, 128, 250, 512, 1024, 2048):
        a = pscenes.scaled_config(n)
        b = common.scaled_config(n)
        assert np.float32(a[0]) == np.float32(b[0]) and a[1] == b[1], (n, a, b)


def test_c_abi_sharding_rule_matches_the_python_harness():
    """pvx_shard_bounds (the rule pvx_multi_solve shards listeners by) == planeverb_b200.sharding.shard_bounds (torchrun harness)"""
    from planeverb_b200 import sharding
    for n in (0, 1, 5, 8, 13):
        for parts in (1, 2, 3, 8):
            covered = []
            for k in range(parts):
                assert pvcuda.shard_bounds(n, parts, k) == sharding.shard_bounds(n, parts, k)
                lo, hi = pvcuda.shard_bounds(n, parts, k)
                covered += list(range(lo, hi))
            assert covered == list(range(n))
