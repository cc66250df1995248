"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI
(include/planeverb_cuda.h, planeverb_ext.h via planeverb_b200/pvcuda.py).  The checker is the oracle:
golden vectors captured from the unmodified reference (tests/golden/) and the plain-C restatement
(oracle/pv_oracle.c), itself pinned bit-for-bit to the reference by tests/test_oracle.py.

Bar: pressure/velocity fields, impulse responses, onset delays, obstruction, wet gain, low-pass cutoff,
source directivity, listener direction AND RT60 BIT-EXACT (the device reproduces glibc's log10f bit for bit,
tests/test_oracle.py::test_device_log10f_recipe_matches_libm); the low-pass cutoff within 1 ulp (powf is
evaluated in double and rounded).  BASELINE.json would allow 1e-4 relative.
"""
import numpy as np
import pytest

from oracle import pvoracle
from tests import common

pytestmark = pytest.mark.gpu

RT60_RTOL = 0.0           # bit-exact (north_star would allow 1e-4)
LOWPASS_RTOL = 2.5e-7     # powf: the device rounds a double pow, libm's powf may differ by 1 ulp on rare cells
EXACT = [0, 1, 4, 5, 6, 7]


@pytest.fixture(scope="module")
def pv():
    from planeverb_b200 import pvcuda
    if pvcuda.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path, GPU tests cannot run")
    return pvcuda


def assert_results(got, got_delay, ref, ref_delay, exclude=None, rt60_rtol=RT60_RTOL):
    assert np.array_equal(got_delay, ref_delay), "onset delays differ"
    valid = ref_delay < 3e38
    if exclude is not None:
        valid &= ~exclude
    every = np.ones_like(valid) if exclude is None else ~exclude
    for k in EXACT:
        m = every if k in (4, 5) else valid
        ok = common.bit_equal(got[m, k], ref[m, k])
        assert ok.all(), f"{common.FIELDS[k]}: {int((~ok).sum())} cells differ (max rel {common.rel_err(got[m, k], ref[m, k]).max():.3e})"
    lp = common.rel_err(got[valid, 3], ref[valid, 3])
    assert lp.size == 0 or lp.max() <= LOWPASS_RTOL, f"lowpass max rel err {lp.max():.3e}"
    # RT60 can legitimately be +-inf / NaN on degenerate regressions (Analyzer.cpp:321-326): compare bits
    same = common.bit_equal(got[valid, 2], ref[valid, 2]) | (np.isnan(got[valid, 2]) & np.isnan(ref[valid, 2]))
    e = common.rel_err(got[valid, 2][~same], ref[valid, 2][~same])
    assert e.size == 0 or e.max() <= rt60_rtol, f"rt60: {e.size} cells differ, max rel err {e.max():.3e} > {rt60_rtol:.1e}"
    return float(e.max()) if e.size else 0.0


def run_pair(pv, scenes, scene, n=None, res=275, T=0, listeners=None, **kw):
    if n is None:
        size, scale = 25.0, 1.0
    else:
        size, scale = common.scaled_config(n, res)
    listeners = listeners or [tuple(v * scale for v in common.DEFAULT_LISTENER)]
    ora = pvoracle.OracleSim(size, size, res, T=T)
    gpu = pv.Scene(size, size, res, T=T, max_sources=len(listeners), **kw)
    assert (gpu.gx, gpu.gy, gpu.T, gpu.fs) == (ora.gx, ora.gy, ora.T, ora.fs)
    assert gpu.efree == ora.efree
    for b in (common.boxes_of(scenes, scene, scale) if scene else []):
        ora.add_aabb(*b)
        gpu.add_aabb(*b)
    return gpu, ora, listeners


@pytest.mark.parametrize("name", common.GOLDEN_CASES)
def test_golden_vectors(pv, name):
    """Outputs of the unmodified reference (tools/make_golden.py) reproduced on the device."""
    meta, z = common.load_golden(name)
    gpu = pv.Scene(meta["size"], meta["size"], meta["resolution"], T=meta["T_override"])
    assert (gpu.gx, gpu.gy, gpu.T, gpu.fs) == (meta["gx"], meta["gy"], meta["T"], meta["fs"])
    dx, dt, efree, courant = z["scalars"]
    assert (gpu.dx, gpu.dt, gpu.efree, gpu.courant) == (dx, dt, efree, courant)
    assert np.array_equal(gpu.pulse(), z["pulse"])
    for b in common.golden_boxes(z):
        gpu.add_aabb(*b)
    gb, gy = gpu.coef()
    assert np.array_equal(gb, z["b"])
    R = z["R"]
    assert common.bit_equal(gy[gb == 0], ((1 - R) / (1 + R))[gb == 0]).all()
    res, dly = gpu.solve([meta["listener"]])
    for k, t in enumerate(meta["snaps_t"]):
        assert common.bit_equal(gpu.pressure(t), z["snap_p"][k]).all(), f"pressure plane t={t}"
    pr, pc = meta["probe"]
    ir = gpu.ir(pr, pc)
    for f, fname in enumerate(("p", "vx", "vy")):
        assert common.bit_equal(ir[:, f], z["ir"][:, f]).all(), f"impulse response {fname}"
    clamped = common.reference_clamped(meta, z["delay"], gpu.D)
    assert_results(res[0], dly[0], z["results"], z["delay"], exclude=clamped)
    gpu.close()


@pytest.mark.parametrize("step_kernel,variant", [(1, 0), (0, 0), (0, 18), (0, 47), (0, 50), (0, 63), (0, 64), (0, 65), (0, 66), (0, 67), (0, 69), (0, 70), (0, 71), (0, 72)])
def test_every_step_kernel_variant_matches_oracle(pv, scenes, step_kernel, variant):
    gpu, ora, Ls = run_pair(pv, scenes, "FloorPlanScene", n=250, T=301, step_kernel=step_kernel, variant=variant)
    res, dly = gpu.solve(Ls)
    ora.generate(Ls[0], keep_velocity=True)
    ora.analyze(Ls[0])
    for t in (0, 1, 2, 3, 4, 5, 150, 299, 300):
        assert common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(gpu.gx + 1, gpu.gy + 1)).all(), f"t={t}"
    p, vx, vy = gpu.state()
    shp = p.shape
    assert common.bit_equal(p, ora.hist[-1].reshape(shp)).all()
    assert common.bit_equal(vx, ora.hvx[-1].reshape(shp)).all()
    assert common.bit_equal(vy, ora.hvy[-1].reshape(shp)).all()
    assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    gpu.close()


@pytest.mark.parametrize("scene,res", [("FloorPlanScene", 750), ("BigRoom", 600)])
def test_contract_mode_at_high_resolutions(pv, scenes, scene, res):
    """The reference's own contract -- the Sandbox's 25 m world with everything derived from the resolution (cfg 1) -- at
    resolutions the golden set does not hold: 750 -> 191 x 191 cells, T = 1187 (the largest the Sandbox offers), and 600.  Grid
    parameters, the free-field normaliser computed on the device, planes and every analyzer output against the oracle."""
    gpu, ora, Ls = run_pair(pv, scenes, scene, res=res)
    if res == 750:
        assert (gpu.gx, gpu.gy, gpu.T) == (191, 191, 1187)
    res_, dly = gpu.solve(Ls)
    ora.generate(Ls[0])
    ora.analyze(Ls[0])
    for t in (0, 5, gpu.T // 2, gpu.T - 1):
        assert common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(gpu.gx + 1, gpu.gy + 1)).all(), f"t={t}"
    assert (ora.delay < 3e38).sum() > 1000
    assert_results(res_[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    gpu.close()


def test_config1_smallroom_128_500(pv, scenes):
    """BASELINE.json configs[0]: SmallRoom.pv, 128x128, 1 source, 500 steps."""
    gpu, ora, Ls = run_pair(pv, scenes, "SmallRoom", n=128, T=500)
    res, dly = gpu.solve(Ls)
    ora.generate(Ls[0]); ora.analyze(Ls[0])
    assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    for (x, z) in common.EMITTERS:
        pos = (x * 128 * float(ora.dx) / 25.0, 0.0, z * 128 * float(ora.dx) / 25.0)
        out = gpu.lookup(pos)
        cell = (int(np.float32(pos[0]) / ora.dx), int(np.float32(pos[2]) / ora.dx))      # Analyzer.cpp:110-111
        assert out is not None
        assert common.bit_equal(out[EXACT], ora.results[cell[0] * gpu.gx + cell[1], EXACT]).all()
        assert common.rel_err(out[2:4], ora.results[cell[0] * gpu.gx + cell[1], 2:4]).max() <= LOWPASS_RTOL
    gpu.close()


def test_config2_shoebox_512_2000(pv, scenes):
    """BASELINE.json configs[1]: Shoebox.pv, 512x512, 1 source, 2000 steps (oracle: the C restatement)."""
    gpu, ora, Ls = run_pair(pv, scenes, "Shoebox", n=512, T=2000)
    res, dly = gpu.solve(Ls)
    ora.generate(Ls[0]); ora.analyze(Ls[0])
    for t in (0, 7, 500, 1999):
        assert common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(gpu.gx + 1, gpu.gy + 1)).all(), f"t={t}"
    worst = assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    print(f"config2 rt60 max rel err {worst:.2e}")
    gpu.close()


def test_batched_sources_equal_separate_solves(pv, scenes):
    """Sources are independent: a batch of 3 listeners == 3 single solves, bit for bit (the multi-GPU
    sharding relies on exactly this)."""
    size, scale = common.scaled_config(300)
    Ls = common.listeners_for(3, scale)
    boxes = common.boxes_of(scenes, "HugeRoom", scale)
    batch = pv.Scene(size, size, 275, T=400, max_sources=3)
    for b in boxes:
        batch.add_aabb(*b)
    rb, db = batch.solve(Ls)
    single = pv.Scene(size, size, 275, T=400, max_sources=1)
    for b in boxes:
        single.add_aabb(*b)
    for i, L in enumerate(Ls):
        single.clear_results(0)
        rs, ds = single.solve([L])
        assert np.array_equal(ds[0], db[i])
        assert np.array_equal(rs[0].view(np.uint32), rb[i].view(np.uint32))
        assert common.bit_equal(single.pressure(399), batch.pressure(399, i)).all()
    batch.close(); single.close()


@pytest.mark.parametrize("n,T,listener_cell", [
    (5, 40, (2, 2)),            # tiny grid, T shorter than the analysis windows (clamped cells excluded)
    (119, 133, (0, 0)),         # listener in the corner cell, T not a multiple of the 4-step block
    (120, 97, (119, 119)),      # grid exactly one tile wide, listener in the last interior cell
    (121, 150, (60, 120)),      # one column past a tile boundary, listener on the last interior column
    (239, 202, (238, 0)),
    (97, 64, (50, 50)),
    (130, 90, (130, 40)),       # listener on the padding ROW (b == 0): records zeros, the injected samples never propagate
    (130, 91, (40, 130)),       # listener on the padding COLUMN
    (60, 50, (60, 60)),         # ... on the padding corner
])
def test_edge_case_grids_and_listeners(pv, n, T, listener_cell):
    size, _ = common.scaled_config(n)
    ora = pvoracle.OracleSim(size, size, 275, T=T, efree=0.0447895788)
    gpu = pv.Scene(size, size, 275, T=T, efree=0.0447895788)
    dx = float(ora.dx)
    L = ((listener_cell[0] + 0.5) * dx, 0.0, (listener_cell[1] + 0.5) * dx)
    assert ora.listener_cell(L) == listener_cell
    # a wall clipped by the grid edge and one covering the padding row/column (Grid.cpp:231,235 clip inclusive)
    for b in [(-0.5 * dx, 0.3 * n * dx, 3 * dx, 4 * dx, 0.9), (n * dx, 0.7 * n * dx, 4 * dx, 6 * dx, 0.5),
              (0.6 * n * dx, n * dx, 5 * dx, 2.5 * dx, 0.97)]:
        ora.add_aabb(*b); gpu.add_aabb(*b)
    res, dly = gpu.solve([L])
    ora.generate(L, keep_velocity=True); ora.analyze(L)
    for t in sorted(set([0, 1, 2, 3, 4, T // 2, T - 1])):
        assert common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(n + 1, n + 1)).all(), f"t={t}"
    p, vx, vy = gpu.state()
    assert common.bit_equal(vx, ora.hvx[-1].reshape(n + 1, n + 1)).all()
    assert common.bit_equal(vy, ora.hvy[-1].reshape(n + 1, n + 1)).all()
    # final pressure state = last record + the last injected sample (FDTD.cpp:234), also when the listener sits on a b == 0 cell
    want_p = ora.hist[T - 1].copy()
    want_p[listener_cell[0] * (n + 1) + listener_cell[1]] += ora.pulse[T - 1]
    assert common.bit_equal(p, want_p.reshape(n + 1, n + 1)).all()
    assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    gpu.close()


@pytest.mark.parametrize("variant,n", [(69, 96), (72, 128), (70, 120), (71, 96), (63, 112), (65, 128), (67, 96)])
def test_resident_tilings_when_the_grid_is_a_multiple_of_the_owned_rows(pv, variant, n):
    """gx a multiple of a tiling's owned rows (24 / 32 / 40 / 48 / 56 / 64 / 8): the padding row is then the first halo row of the
    last tile row and is live state there (pvc_step_res.cu, liveRow).  Listener in the last interior cell, walls on the padding
    row and column, T not a multiple of 4; planes, final state and every output against the oracle, for each tiling explicitly
    (the automatic selection only ever picks one of them per grid size)."""
    T = 97
    size, _ = common.scaled_config(n)
    ora = pvoracle.OracleSim(size, size, 275, T=T, efree=0.0447895788)
    gpu = pv.Scene(size, size, 275, T=T, efree=0.0447895788, variant=variant)
    assert gpu.step_variant() == variant
    dx = float(ora.dx)
    L = ((n - 0.5) * dx, 0.0, (n - 0.5) * dx)
    assert ora.listener_cell(L) == (n - 1, n - 1)
    for b in [(n * dx, 0.5 * n * dx, 4 * dx, 6 * dx, 0.5), (0.4 * n * dx, n * dx, 5 * dx, 2.5 * dx, 0.97), (0.5 * n * dx, 0.5 * n * dx, 7 * dx, 3 * dx, 0.8)]:
        ora.add_aabb(*b); gpu.add_aabb(*b)
    res, dly = gpu.solve([L])
    ora.generate(L, keep_velocity=True); ora.analyze(L)
    for t in (0, 1, 4, 5, T // 2, T - 1):
        assert common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(n + 1, n + 1)).all(), f"t={t}"
    p, vx, vy = gpu.state()
    assert common.bit_equal(vx, ora.hvx[-1].reshape(n + 1, n + 1)).all()
    assert common.bit_equal(vy, ora.hvy[-1].reshape(n + 1, n + 1)).all()
    assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    gpu.close()


def test_listener_inside_a_wall_produces_no_onsets(pv):
    gpu = pv.Scene(25.0, 25.0, 275)
    ora = pvoracle.OracleSim(25.0, 25.0, 275)
    box = (5.0, 4.0, 3.0, 3.0, 0.9)
    gpu.add_aabb(*box); ora.add_aabb(*box)
    L = common.DEFAULT_LISTENER
    res, dly = gpu.solve([L])
    ora.generate(L); ora.analyze(L)
    assert (dly[0] > 3e38).all() and np.array_equal(dly[0], ora.delay)
    assert (res[0][:, [0, 1, 2, 3, 6, 7]] == 0).all()
    assert common.bit_equal(res[0][:, 4:6], ora.results[:, 4:6]).all()      # direction is still written for every cell
    gpu.close()


def test_stale_results_and_geometry_edits_follow_the_reference(pv, scenes):
    """Frame 1: listener A. Then UpdateGeometry (= remove old + add new, GeometryManager.cpp:112-121) and
    frame 2 with listener B WITHOUT clearing: cells with no onset keep frame 1's values (Analyzer.cpp:161-165)."""
    gpu = pv.Scene(25.0, 25.0, 275, T=60)                    # 60 steps: each frame reaches only ~35 cells
    ora = pvoracle.OracleSim(25.0, 25.0, 275, T=60)
    boxes = common.boxes_of(scenes, "SingleWall")
    for b in boxes:
        gpu.add_aabb(*b); ora.add_aabb(*b)
    A, B = (5.0, 0.0, 5.0), (20.0, 0.0, 18.0)
    res1, dly1 = gpu.solve([A])
    ora.generate(A); ora.analyze(A)
    assert_results(res1[0], dly1[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    old, new = boxes[0], (boxes[0][0] + 2.0, boxes[0][1] + 1.0, boxes[0][2], boxes[0][3], 0.8)
    gpu.remove_aabb(*old); gpu.add_aabb(*new)
    ora.remove_aabb(*old); ora.add_aabb(*new)
    gb, _ = gpu.coef()
    assert np.array_equal(gb, ora.coef()[0])
    res2, dly2 = gpu.solve([B])
    ora.generate(B); ora.analyze(B)
    stale = ora.delay > 3e38
    assert stale.any() and (ora.results[stale, 0] != 0).any(), "test needs stale non-zero cells"
    assert np.array_equal(dly2[0], ora.delay)
    for k in EXACT:
        assert common.bit_equal(res2[0][:, k], ora.results[:, k]).all(), common.FIELDS[k]
    ok2 = common.bit_equal(res2[0][:, 2], ora.results[:, 2]) | (np.isnan(res2[0][:, 2]) & np.isnan(ora.results[:, 2]))
    assert ok2.all()
    assert common.rel_err(res2[0][:, 3], ora.results[:, 3]).max() <= LOWPASS_RTOL
    # removing everything returns to the empty-grid solution
    for b in boxes[1:] + [new]:
        gpu.remove_aabb(*b)
    gpu.clear_results(0)
    res3, dly3 = gpu.solve([B])
    empty = pv.Scene(25.0, 25.0, 275, T=60, efree=float(gpu.efree))
    res4, dly4 = empty.solve([B])
    assert np.array_equal(dly3, dly4) and np.array_equal(res3.view(np.uint32), res4.view(np.uint32))
    gpu.close(); empty.close()


def test_full_size_fused_equals_baseline_kernel(pv, scenes):
    """BASELINE.json configs[2] grid and step count (1024x1024, 4000 steps), one source: the temporally blocked
    register-tiled kernel and the two-launch baseline kernel are independent device formulations and must
    agree bit for bit on fields and on every analyzer output; plus determinism of a second run."""
    size, scale = common.scaled_config(1024)
    L = [common.listeners_for(1, scale)[0]]
    boxes = common.boxes_of(scenes, "BigRoom", scale)
    out = []
    for sk in (0, 1):
        sc = pv.Scene(size, size, 275, T=4000, max_sources=1, step_kernel=sk, efree=0.0447895788)
        for b in boxes:
            sc.add_aabb(*b)
        res, dly = sc.solve(L)
        planes = [sc.pressure(t) for t in (3, 1000, 3999)]
        state = sc.state()
        if sk == 0:
            res_again, dly_again = sc.solve(L)
            assert np.array_equal(res.view(np.uint32), res_again.view(np.uint32)) and np.array_equal(dly, dly_again)
        out.append((res, dly, planes, state))
        sc.close()
    (r0, d0, p0, s0), (r1, d1, p1, s1) = out
    assert np.array_equal(d0, d1)
    assert (d0 < 3e38).sum() > 0.1 * d0.size          # BigRoom is a closed room: the wave never leaves it
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
    for a, b in zip(p0 + list(s0), p1 + list(s1)):
        assert common.bit_equal(a, b).all()                 # identical up to the sign of exact zeros
    # physical sanity at full size: outputs finite where an onset exists, unit vectors are unit
    valid = d0[0] < 3e38
    assert np.isfinite(r0[0][valid][:, [0, 1, 3]]).all()
    norm = np.hypot(r0[0][valid, 6], r0[0][valid, 7])
    assert np.allclose(norm[norm > 0], 1.0, atol=1e-5)


def test_rt60_tolerance_at_long_response(pv, scenes):
    """RT60 is the only output that is not bit-exact; check it at a long response (T=4000) against the oracle."""
    gpu, ora, Ls = run_pair(pv, scenes, "Shoebox", n=200, T=4000)
    res, dly = gpu.solve(Ls)
    ora.generate(Ls[0]); ora.analyze(Ls[0])
    worst = assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    print(f"T=4000 rt60 max rel err {worst:.2e}")
    gpu.close()


def test_config4_hugeroom_2048_properties(pv, scenes):
    """BASELINE.json configs[3] grid (2048x2048, HugeRoom.pv) at a bounded step count: the default (auto -> TMA
    persistent) kernel against the two-launch baseline kernel bit for bit, two batched sources == their single
    solves, and a CPU-oracle check of the first 120 steps on a 2048-wide pressure plane."""
    size, scale = common.scaled_config(2048)
    boxes = common.boxes_of(scenes, "HugeRoom", scale)
    Ls = common.listeners_for(2, scale)
    T = 600
    fused = pv.Scene(size, size, 275, T=T, max_sources=2, efree=0.0447895788)
    base = pv.Scene(size, size, 275, T=T, max_sources=1, step_kernel=1, efree=0.0447895788)
    for b in boxes:
        fused.add_aabb(*b); base.add_aabb(*b)
    rf, df = fused.solve(Ls)
    for i, L in enumerate(Ls):
        base.clear_results(0)
        rb, db = base.solve([L])
        assert np.array_equal(db[0], df[i])
        assert np.array_equal(rb[0].view(np.uint32), rf[i].view(np.uint32))
        for t in (0, 3, 4, 299, T - 1):
            assert common.bit_equal(base.pressure(t), fused.pressure(t, i)).all(), (i, t)
    assert (df[0] < 3e38).sum() > 1000
    fused.close(); base.close()
    # oracle on the same 2048 grid, few steps (the pulse has travelled ~80 cells)
    ora = pvoracle.OracleSim(size, size, 275, T=120, efree=0.0447895788)
    gpu = pv.Scene(size, size, 275, T=120, max_sources=1, efree=0.0447895788)
    for b in boxes:
        ora.add_aabb(*b); gpu.add_aabb(*b)
    res, dly = gpu.solve([Ls[0]])
    ora.generate(Ls[0]); ora.analyze(Ls[0])
    for t in (0, 5, 60, 119):
        assert common.bit_equal(gpu.pressure(t), ora.hist[t].reshape(2049, 2049)).all(), t
    assert_results(res[0], dly[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    gpu.close()


def test_config5_dynamic_geometry_frames(pv, scenes):
    """BASELINE.json configs[4] pattern: FloorPlanScene.pv with one AABB moved every frame (UpdateGeometry =
    Remove(old) + Add(new) re-voxelised on the device) followed by a full solve; every frame must equal a fresh
    scene built directly in that configuration, and the oracle on the last frame."""
    n, T = 256, 300
    size, scale = common.scaled_config(n)
    boxes = common.boxes_of(scenes, "FloorPlanScene", scale)
    L = common.listeners_for(1, scale)
    live = pv.Scene(size, size, 275, T=T, efree=0.0447895788)
    for b in boxes:
        live.add_aabb(*b)
    door = boxes[3]
    cur = door
    for frame in range(4):
        moved = (door[0] + 0.4 * (frame + 1) * scale, door[1] - 0.3 * (frame + 1) * scale, door[2], door[3], door[4])
        live.remove_aabb(*cur); live.add_aabb(*moved)
        cur = moved
        live.clear_results(0)
        r_live, d_live = live.solve(L)
        fresh = pv.Scene(size, size, 275, T=T, efree=0.0447895788)
        ora = pvoracle.OracleSim(size, size, 275, T=T, efree=0.0447895788)
        for b in boxes:
            fresh.add_aabb(*b); ora.add_aabb(*b)
        c2 = door
        for k in range(frame + 1):      # the reference has no overlap ref-counting: replay the same edit history
            m2 = (door[0] + 0.4 * (k + 1) * scale, door[1] - 0.3 * (k + 1) * scale, door[2], door[3], door[4])
            fresh.remove_aabb(*c2); fresh.add_aabb(*m2)
            ora.remove_aabb(*c2); ora.add_aabb(*m2)
            c2 = m2
        assert np.array_equal(live.coef()[0], ora.coef()[0])
        r_fresh, d_fresh = fresh.solve(L)
        assert np.array_equal(d_live, d_fresh) and np.array_equal(r_live.view(np.uint32), r_fresh.view(np.uint32))
        fresh.close()
    ora.generate(L[0]); ora.analyze(L[0])
    assert_results(r_live[0], d_live[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    live.close()


def test_pipelined_frame_loop_equals_synchronous_solves(pv, scenes):
    """pvx_solve_pipelined / pvx_fetch_wait (the frame-loop form: the D2H copy of frame k's result grids overlaps the time
    steps of frame k+1, geometry edits and listeners staged through pinned rings so the host never waits for the stream):
    every frame's host grids must equal, bit for bit, a synchronous solve of a fresh scene in the same state -- including
    the stale values cells without an onset keep from the frame before (Analyzer.cpp:161-165)."""
    n, T, S = 256, 300, 2
    size, scale = common.scaled_config(n)
    boxes = common.boxes_of(scenes, "FloorPlanScene", scale)
    door = boxes[3]
    cells = n * n
    live = pv.Scene(size, size, 275, T=T, max_sources=S, efree=0.0447895788)
    sync = pv.Scene(size, size, 275, T=T, max_sources=S, efree=0.0447895788)
    for b in boxes:
        live.add_aabb(*b); sync.add_aabb(*b)
    bufs = [(pv.pinned_array((S, cells, 8)), pv.pinned_array((S, cells))) for _ in range(2)]
    frames = 5
    expect = []
    cur = door
    for k in range(frames):
        moved = (door[0] + 0.4 * (k + 1) * scale, door[1] - 0.3 * (k + 1) * scale, door[2], door[3], door[4])
        L = common.listeners_for(S, scale)
        L = [(x + 0.2 * k * scale, y, z) for (x, y, z) in L]
        sync.remove_aabb(*cur); sync.add_aabb(*moved)
        r, d = sync.solve(L)
        expect.append((r.copy(), d.copy()))
        live.remove_aabb(*cur); live.add_aabb(*moved)
        live.solve_pipelined(L, bufs[k & 1])                    # returns once frame k-1's grids are on the host
        if k > 0:
            r0, d0 = bufs[(k - 1) & 1]
            assert np.array_equal(d0, expect[k - 1][1]), f"frame {k - 1}: delays differ"
            assert np.array_equal(r0.view(np.uint32), expect[k - 1][0].view(np.uint32)), f"frame {k - 1}: results differ"
        cur = moved
    live.fetch_wait()
    r0, d0 = bufs[(frames - 1) & 1]
    assert np.array_equal(d0, expect[-1][1]) and np.array_equal(r0.view(np.uint32), expect[-1][0].view(np.uint32))
    live.fetch_wait()                                           # idempotent
    # pipelined per-emitter lookups of the same frame: equal to the synchronous lookup, -1 outside the grid
    ems = [(5.0 * scale, 0.0, 6.0 * scale), (12.5 * scale, 0.0, 12.5 * scale), (-3.0, 0.0, 1.0), (20.0 * scale, 0.0, 20.0 * scale)]
    eb = pv.pinned_array((S, len(ems), 8))
    live.lookup_wait(live.lookup_async(ems, eb))
    for s_ in range(S):
        for e, pos in enumerate(ems):
            ref = live.lookup(pos, s_)
            if ref is None:
                assert (eb[s_, e] == -1.0).all()
            else:
                assert np.array_equal(eb[s_, e].view(np.uint32), ref.view(np.uint32))
    # a synchronous call after the pipelined ones still works on the same scene
    r1, d1 = live.solve(L)
    assert np.array_equal(r1.view(np.uint32), expect[-1][0].view(np.uint32)) and np.array_equal(d1, expect[-1][1])
    live.close(); sync.close()


def test_non_square_grid_is_consistent(pv):
    """The reference mixes strides on non-square grids; this library indexes them consistently:
    fields must match the oracle's solver (which uses the FDTD stride gy+1 throughout), the fused and baseline
    kernels must agree on every analyzer output, and lookups must address row*gy + col."""
    res = 275
    dx = float(pvoracle.grid_params(res)[0])
    sx, sy = (150 + 0.5) * dx, (97 + 0.5) * dx
    T = 260
    ora = pvoracle.OracleSim(sx, sy, res, T=T, efree=0.0447895788)
    assert (ora.gx, ora.gy) == (150, 97)
    out = []
    for sk in (0, 1):
        g = pv.Scene(sx, sy, res, T=T, efree=0.0447895788, step_kernel=sk)
        assert (g.gx, g.gy) == (150, 97)
        for b in [(20 * dx, 40 * dx, 30 * dx, 3 * dx, 0.9), (100 * dx, 60 * dx, 4 * dx, 50 * dx, 0.7)]:
            g.add_aabb(*b)
            if sk == 0:
                ora.b.reshape(151, 98)          # (oracle AddAABB assumes square strides: set its fields from the device below)
        L = (75.2 * dx, 0.0, 30.7 * dx)
        r, d = g.solve([L])
        out.append((r, d, g.pressure(T - 1), g.state(), g.coef()))
        probe = g.lookup((10.5 * dx, 0.0, 90.5 * dx))
        assert probe is not None and np.array_equal(probe.view(np.uint32), r[0][10 * 97 + 90].view(np.uint32))
        assert g.lookup((10.5 * dx, 0.0, 97.5 * dx)) is None and g.lookup((150.5 * dx, 0.0, 5 * dx)) is None
        g.close()
    (r0, d0, p0, s0, c0), (r1, d1, p1, s1, c1) = out
    assert np.array_equal(d0, d1) and np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
    assert common.bit_equal(p0, p1).all() and all(common.bit_equal(a, b).all() for a, b in zip(s0, s1))
    # oracle field check with the device's wall plane copied in (b, R such that (1-R)/(1+R) = Y)
    b_dev, y_dev = c0
    ora.b[:] = b_dev.reshape(-1)
    ora.R[:] = np.where(b_dev.reshape(-1) == 0, (1 - y_dev.reshape(-1)) / (1 + y_dev.reshape(-1)), 0).astype(np.float32)
    ora.generate(L)
    # admittances re-derived from R may differ by an ulp from the device's Y, so compare an air-only early plane exactly
    # and the late plane to 1e-6
    assert np.allclose(p0, ora.hist[T - 1].reshape(151, 98), rtol=0, atol=2e-7)


def test_error_paths(pv):
    with pytest.raises(pv.PlaneverbCudaError):
        pv.Scene(25.0, 25.0, 275, device=99)
    sc = pv.Scene(25.0, 25.0, 275, max_sources=2)
    with pytest.raises(pv.PlaneverbCudaError):
        sc.solve([(5, 0, 4)] * 3)                      # more listeners than max_sources
    with pytest.raises(pv.PlaneverbCudaError):
        sc.solve([(500.0, 0, 4)])                      # listener outside the grid
    with pytest.raises(pv.PlaneverbCudaError):
        sc.pressure(10 ** 6)
    assert sc.lookup((-3.0, 0, 1.0)) is None
    res, dly = sc.solve([(5, 0, 4)])                    # still usable after errors
    assert (dly[0] < 3e38).any()
    sc.close()
    with pytest.raises(pv.PlaneverbCudaError) as e:
        pv.Scene(25.0, 25.0, 275, T=10)                # free-field probe needs 18 samples (FreeGrid.cpp:100)
    assert "invalid" in str(e.value)


def test_outputs_through_the_reference_dsp_consumer(pv, scenes):
    """the consumer of the hot path's outputs is PlaneverbDSP.  The UNMODIFIED reference consumer
    (PlaneverbDSP/src/PvDSPContext.cpp, built into oracle/_ref/libpvdspref.so by oracle/dspdriver/Makefile) renders a test
    signal with the device's PlaneverbOutput of every sampled cell and with the reference's own golden output of the same
    cell: its validity gates (lowpass in [20, 20000] Hz, obstructionGain > 0, direction != 0, :258-262) must decide alike, and
    the dry bus and the three reverb buses (FindGainA/B/C, :165-229; Butterworth low-pass, Lowpass.cpp) must come out
    bit-identical wherever the low-pass cutoffs are (the cutoff may differ by 1 ulp: powf), else within 1e-5."""
    from oracle import pvdspref
    assert pvdspref.available(), "oracle/_ref/libpvdspref.so missing: run __graft_entry__.build() where /root/reference exists"
    audio = pvdspref.test_signal(256)
    for name in ("singlewall_95_res375", "floorplan_70"):
        meta, z = common.load_golden(name)
        gpu = pv.Scene(meta["size"], meta["size"], meta["resolution"], T=meta["T_override"])
        for b in common.golden_boxes(z):
            gpu.add_aabb(*b)
        res, dly = gpu.solve([meta["listener"]])
        valid = (dly[0] < 3e38) & ~common.reference_clamped(meta, z["delay"], gpu.D)
        cells = np.nonzero(valid)[0]
        cells = cells[:: max(1, cells.size // 400)]
        lx, lz = meta["listener"][0], meta["listener"][2]
        accepted = exact = 0
        for cell in cells:
            r, c = divmod(int(cell), gpu.gy)
            pos = ((r + 0.5) * float(gpu.dx), (c + 0.5) * float(gpu.dx))
            out = gpu.lookup((pos[0], 0.0, pos[1]))                       # Analyzer::GetResponseResult through the C ABI
            assert out is not None and common.bit_equal(out, res[0][cell]).all()
            dev = pvdspref.render(out, pos, (lx, lz), audio)
            ref = pvdspref.render(z["results"][cell], pos, (lx, lz), audio)
            assert (np.abs(dev).sum() > 0) == (np.abs(ref).sum() > 0), f"{name} cell {cell}: the DSP gates disagree"
            if np.abs(ref).sum() > 0:
                accepted += 1
                if out[3] == z["results"][cell][3]:
                    assert np.array_equal(dev.view(np.uint32), ref.view(np.uint32)), f"{name} cell {cell}: DSP buses differ"
                    exact += 1
                else:
                    assert np.allclose(dev, ref, rtol=1e-5, atol=1e-7), f"{name} cell {cell}: DSP buses differ beyond the 1-ulp cutoff"
        assert accepted > 0.9 * cells.size and exact > 0.9 * accepted, (name, cells.size, accepted, exact)
        gpu.close()


@pytest.mark.parametrize("scene,n,T", [("HugeRoom", 400, 900), ("FloorPlanScene", 300, 700), (None, 257, 600)])
def test_pointer_jumping_direction_equals_the_sequential_walk(pv, scenes, scene, n, T):
    """Analyzer::EncodeListenerDirection (Analyzer.cpp:340-431): the default device path resolves every walk by
    pointer jumping over a link array; pvc_set_walk_mode(1) runs the reference's walk one thread per cell.  Both must
    give the same direction for every cell (and the oracle agrees, see the other parity tests)."""
    size, scale = common.scaled_config(n)
    listeners = common.listeners_for(2, scale)
    outs = []
    for sequential in (False, True):
        gpu = pv.Scene(size, size, 275, T=T, max_sources=2)
        gpu.set_walk_mode(sequential)
        for b in (common.boxes_of(scenes, scene, scale) if scene else []):
            gpu.add_aabb(*b)
        res, dly = gpu.solve(listeners)
        outs.append((np.array(res, copy=True), np.array(dly, copy=True)))
        gpu.close()
    (ra, da), (rb, db) = outs
    assert np.array_equal(da, db)
    assert common.bit_equal(ra, rb).all()
    assert (np.abs(ra[..., 4:6]).sum(axis=-1) > 0).mean() > 0.5      # the directions are not trivially zero
