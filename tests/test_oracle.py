"""CPU tests (no GPU): the oracle's C restatement against (a) the golden vectors captured from the
unmodified reference and (b) the reference itself when oracle/_ref is present, plus known-answer
checks derived from the reference's own formulas."""
import numpy as np
import pytest

from oracle import pvoracle, pvref
from tests import common


def _oracle_for(meta, z, keep_velocity=False):
    sim = pvoracle.OracleSim(meta["size"], meta["size"], meta["resolution"], T=meta["T_override"])
    for b in common.golden_boxes(z):
        sim.add_aabb(*b)
    sim.generate(meta["listener"], keep_velocity=keep_velocity)
    sim.analyze(meta["listener"])
    return sim


@pytest.mark.parametrize("name", common.GOLDEN_CASES)
def test_oracle_matches_golden_bit_for_bit(name):
    meta, z = common.load_golden(name)
    sim = _oracle_for(meta, z, keep_velocity=True)
    assert (sim.gx, sim.gy, sim.T, sim.fs) == (meta["gx"], meta["gy"], meta["T"], meta["fs"])
    dx, dt, efree, courant = z["scalars"]
    assert (sim.dx, sim.dt, sim.efree, sim.courant) == (dx, dt, efree, courant)
    assert np.array_equal(sim.pulse, z["pulse"])
    b, R = sim.coef()
    assert np.array_equal(b, z["b"]) and np.array_equal(R, z["R"])
    for k, t in enumerate(meta["snaps_t"]):
        p, vx, vy = sim.snapshot(t)
        assert common.bit_equal(p, z["snap_p"][k]).all(), f"p plane t={t}"
        assert common.bit_equal(vx, z["snap_vx"][k]).all(), f"vx plane t={t}"
        assert common.bit_equal(vy, z["snap_vy"][k]).all(), f"vy plane t={t}"
    pr, pc = meta["probe"]
    i = pr * (sim.gy + 1) + pc
    assert common.bit_equal(sim.hist[:, i], z["ir"][:, 0]).all()
    assert common.bit_equal(sim.hvx[:, i], z["ir"][:, 1]).all()
    assert common.bit_equal(sim.hvy[:, i], z["ir"][:, 2]).all()
    assert np.array_equal(sim.delay, z["delay"])
    ok = (z["delay"] < 3e38) & (sim.clamped == 0)
    for k, fname in enumerate(common.FIELDS):
        m = np.ones_like(ok) if k in (4, 5) else ok
        assert common.bit_equal(sim.results[m, k], z["results"][m, k]).all(), fname


@pytest.mark.skipif(not pvref.available(), reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("scene,res,listener", [("ExampleProject", 275, (7.0, 0, 9.5)), ("UnityReplicationTest", 375, (3.0, 0, 3.0)),
                                               ("SmallRoomScene", 275, (5.0, 0, 5.0))])
def test_oracle_matches_live_reference(scenes, scene, res, listener):
    ref = pvref.RefSim(25.0, 25.0, res)
    ora = pvoracle.OracleSim(25.0, 25.0, res)
    assert ref.pulse_mismatch == 0
    assert (ref.gx, ref.T, ref.fs, ref.dx, ref.efree, ref.courant) == (ora.gx, ora.T, ora.fs, ora.dx, ora.efree, ora.courant)
    for b in common.boxes_of(scenes, scene):
        ref.add_aabb(*b)
        ora.add_aabb(*b)
    # exercise RemoveAABB / re-add in queue order (GeometryManager.cpp:112-121)
    first = common.boxes_of(scenes, scene)[0]
    moved = (first[0] + 1.3, first[1] - 0.7, first[2], first[3], first[4])
    ref.remove_aabb(*first); ora.remove_aabb(*first)
    ref.add_aabb(*moved); ora.add_aabb(*moved)
    rb, rR = ref.coef()
    ob, oR = ora.coef()
    assert np.array_equal(rb, ob) and np.array_equal(rR, oR)
    ref.generate(listener); ora.generate(listener)
    for t in (0, 3, 40, ref.T - 1):
        assert common.bit_equal(ref.snapshot(t)[0], ora.snapshot(t)[0]).all()
    ref.analyze(listener); ora.analyze(listener)
    rres, rdelay = ref.results()
    assert np.array_equal(rdelay, ora.delay)
    ok = (rdelay < 3e38) & (ora.clamped == 0)
    for k, fname in enumerate(common.FIELDS):
        m = np.ones_like(ok) if k in (4, 5) else ok
        assert common.bit_equal(ora.results[m, k], rres[m, k]).all(), fname
    # the reference's own emitter lookup agrees with indexing our grid
    for (x, zz) in common.EMITTERS:
        r = ref.lookup((x, 0, zz))
        cell = int(np.float32(x) / ora.dx) * ora.gx + int(np.float32(zz) / ora.dx)
        assert r is not None and common.bit_equal(r[:4], ora.results[cell, :4]).all()


def test_grid_parameter_table():
    """SURVEY.md App. A table, derived from Grid.cpp:390-396 / :55 / Analyzer.cpp:170-171,237,284."""
    table = {275: (1443, 435, (7, 14, 115, 14), 70), 375: (1968, 593, (9, 19, 157, 19), 95),
             500: (2625, 791, (13, 26, 210, 26), 127), 750: (3937, 1187, (19, 39, 314, 39), 191)}
    for res, (fs, T, win, cells) in table.items():
        dx, dt, f = pvoracle.grid_params(res)
        gx, gy, TT, courant = pvoracle.derived(res, 25.0, 25.0)
        assert (f, TT, gx, gy) == (fs, T, cells, cells)
        assert pvoracle.windows(f) == win
        assert abs(float(courant) - 2.0 / 3.0) < 1e-6
        assert abs(float(dx) - (343.21 / res) / 3.5) < 1e-6


def test_gaussian_pulse_known_answer():
    """exp(-(t-2s)^2/s^2), s = 1/(0.5*pi*res): peak ~1 near sample 2*s*fs ~ 6.68 for every resolution."""
    for res in (275, 375, 500, 750):
        _, _, fs = pvoracle.grid_params(res)
        p = pvoracle.gaussian_pulse(res, fs, 64)
        sigma = 1.0 / (0.5 * np.pi * res)
        t = np.arange(64) / fs
        want = np.exp(-((t - 2 * sigma) ** 2) / sigma ** 2)
        assert np.allclose(p, want, rtol=2e-4, atol=1e-7)
        assert int(np.argmax(p)) == 7 and p[40:].max() < 1e-30


def test_free_field_obstruction_is_about_one():
    """EFree/r normalisation (FreeGrid.cpp:57-58,88-91): with no geometry the obstruction gain of cells a
    few metres from the listener is ~1."""
    sim = pvoracle.OracleSim(25.0, 25.0, 275)
    L = (12.5, 0.0, 12.5)
    sim.generate(L); sim.analyze(L)
    res = sim.results.reshape(sim.gx, sim.gy, 8)
    lr, lc = sim.listener_cell(L)
    ring = [res[lr + 10, lc, 0], res[lr, lc + 10, 0], res[lr - 10, lc, 0], res[lr, lc - 10, 0], res[lr + 3, lc, 0]]
    assert all(0.85 < v < 1.15 for v in ring), ring
    # source directivity of a free-field cell points back at the listener (negated flux)
    sd = res[lr + 10, lc, 6:8]
    assert sd[0] < -0.99 and abs(sd[1]) < 0.1


def test_rt60_known_answer_exponential_decay():
    """SchroederEnvelope.sci recipe: an exponentially decaying IR with a known T60 must come back from the
    backward-integration + regression estimator (Analyzer.cpp:282-326)."""
    import ctypes as C
    fs, T, gx = 1443, 1400, 2
    S = gx + 1
    N = S * S
    t60 = 0.6
    tt = np.arange(T) / fs
    rng = np.random.default_rng(7)
    sig = (10 ** (-3 * tt / t60) * rng.standard_normal(T)).astype(np.float32)     # -60 dB at t60
    hist = np.zeros((T, N), np.float32)
    hist[:, 0] = sig
    onset = np.full(N, -1, np.int32); onset[0] = 0
    edry = np.ones(N, np.float32); fx = np.ones(N, np.float32); fy = np.zeros(N, np.float32); wet = np.ones(N, np.float32)
    results = np.zeros((gx * gx, 8), np.float32); delay = np.zeros(gx * gx, np.float32); clamped = np.zeros(gx * gx, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pvoracle.lib().pvo_encode(gx, gx, T, fs, np.float32(0.3566), np.float32(0.0448), 0.0, 0.0, p(hist), p(onset), p(edry), p(fx), p(fy),
                              p(wet), p(results), p(delay), p(clamped))
    assert abs(results[0, 2] - t60) / t60 < 0.05, results[0, 2]
    assert delay[0] == 0 and delay[1] > 3e38


def test_no_onset_cells_keep_stale_results():
    """Analyzer.cpp:161-165: a cell without an onset is left untouched."""
    sim = pvoracle.OracleSim(25.0, 25.0, 275, T=60)       # wave reaches only ~40 cells in 60 steps
    sim.results[:] = 123.0
    L = (2.0, 0.0, 2.0)
    sim.generate(L); sim.analyze(L)
    far = (sim.gx - 1) * sim.gx + (sim.gy - 1)
    assert sim.delay[far] > 3e38 and (sim.results[far, [0, 1, 2, 3, 6, 7]] == 123.0).all()


def test_device_log10f_recipe_matches_libm():
    """planeverb_b200/csrc/pvc_analyze.cu reproduces glibc's log10f (fdlibm wrapper around the table-driven
    double-precision logf kernel) so that RT60 is bit-exact.  Re-evaluate that recipe here in numpy, with the
    table constants parsed out of the .cu source, and compare with the libm this oracle links."""
    import ctypes as C
    import os
    import re
    src = open(os.path.join(common.ROOT, "planeverb_b200", "csrc", "pvc_analyze.cu")).read()
    body = src[src.index("kLogfTable[16]"):]
    body = body[:body.index("};")]
    vals = [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[+-]\d+", body)]
    assert len(vals) == 32
    tab = np.array(vals).reshape(16, 2)
    ln2 = float.fromhex("0x1.62e42fefa39efp-1")
    A = [float.fromhex(v) for v in ("-0x1.00ea348b88334p-2", "0x1.5575b0be00b6ap-2", "-0x1.ffffef20a4123p-2")]
    for v in ("0x1.62e42fefa39efp-1", "-0x1.00ea348b88334p-2", "0x1.5575b0be00b6ap-2", "-0x1.ffffef20a4123p-2"):
        assert v in src

    def logf(x):
        ix = x.view(np.uint32).astype(np.int64)
        tmp = (ix - 0x3f330000) & 0xffffffff
        i = (tmp >> 19) & 15
        k = tmp.astype(np.uint32).view(np.int32) >> 23
        iz = (ix - (tmp & 0xff800000)) & 0xffffffff
        z = iz.astype(np.uint32).view(np.float32).astype(np.float64)
        r = z * tab[i, 0] - 1.0
        y0 = tab[i, 1] + k.astype(np.float64) * ln2
        r2 = r * r
        y = A[1] * r + A[2]
        y = A[0] * r2 + y
        y = y * r2 + (y0 + r)
        out = y.astype(np.float32)
        out[x == np.float32(1.0)] = 0.0
        return out

    def log10f(x):
        f32 = np.float32
        hx = x.view(np.int32).astype(np.int64)
        k = (hx >> 23) - 127
        i = (k < 0).astype(np.int64)
        m = ((hx & 0x007fffff) | ((0x7f - i) << 23)).astype(np.uint32).view(np.float32)
        y = (k + i).astype(np.float32)
        z = (y * f32(7.9034151668e-07)).astype(f32) + (f32(4.3429449201e-01) * logf(m)).astype(f32)
        return (z.astype(f32) + (y * f32(3.0102920532e-01)).astype(f32)).astype(f32)

    libm = C.CDLL("libm.so.6")
    libm.log10f.restype = C.c_float
    libm.log10f.argtypes = [C.c_float]
    rng = np.random.default_rng(3)
    x = np.concatenate([(10.0 ** rng.uniform(-37, 4, 60000)).astype(np.float32),
                        np.float32([1.0, 0.5, 2.0, 10.0, 1e-30, 0.99999994, 1.0000001])])
    want = np.array([libm.log10f(float(v)) for v in x], np.float32)
    got = log10f(x)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_device_logf_fma_form_is_exhaustively_exact(tmp_path):
    """The analyzer's hot loop evaluates glibc's logf kernel in FMA form over a 33-entry {invc * 2^-k, logc + k*ln2} table
    (pvc_analyze.cu::decibelsNormal).  tools/micro/logf_fma_recipe.c restates that form in C and runs ALL 2^24 floats of
    [0.5, 2) -- the only inputs fdlibm's log10f hands to logf -- against the host libm, together with e_logf.c as written
    and its FMA-contracted variant: every one of them must agree bit for bit on every input."""
    import os
    import re
    import subprocess
    csrc = os.path.join(common.ROOT, "tools", "micro", "logf_fma_recipe.c")
    exe = str(tmp_path / "logf_fma_recipe")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, csrc, "-lm"])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "inputs 16777216  a!=b 0  a!=libm 0  b!=libm 0  device!=libm 0  horner!=libm 0" in p.stdout
    # the C restatement and the CUDA source use the same constants and the same table indexing
    cu = open(os.path.join(common.ROOT, "planeverb_b200", "csrc", "pvc_analyze.cu")).read()
    c = open(csrc).read()
    for const in re.findall(r"-?0x1\.[0-9a-f]+p[+-]\d+", c):
        assert const in cu, const
    assert "0x3f330000u) >> 19) + kLogf33Bias" in cu and "kLogf33Bias = 7" in cu and ">> 19) + 7" in c
    # ... and the device evaluates the Horner chain the harness calls (d): A0, A1, A2, 1, logc in that order
    body = cu[cu.index("decibelsNormal(float e"):]
    body = body[:body.index("#endif")]
    order = [body.index(t) for t in ("-0x1.00ea348b88334p-2, r, 0x1.5575b0be00b6ap-2", "q, r, -0x1.ffffef20a4123p-2", "q, r, 1.0", "q, r, en.logc")]
    assert order == sorted(order)


def test_reference_dsp_consumer_accepts_the_golden_outputs():
    """the unmodified PlaneverbDSP (oracle/_ref/libpvdspref.so) renders the reference's golden outputs: cells
    with an onset pass its validity gates (PvDSPContext.cpp:258-262) and drive the dry bus and, by RT60 (:165-229), the
    matching reverb buses; an all-zero PlaneverbOutput (a cell without an onset) is rejected and leaves silence."""
    from oracle import pvdspref
    if not pvdspref.available():
        pytest.skip("oracle/_ref/libpvdspref.so not built (needs /root/reference)")
    meta, z = common.load_golden("singlewall_95_res375")
    audio = pvdspref.test_signal(256)
    valid = np.nonzero(z["delay"] < 3e38)[0]
    lx, lz = meta["listener"][0], meta["listener"][2]
    n_ok = 0
    for cell in valid[::97]:
        out = z["results"][cell]
        bufs = pvdspref.render(out, (3.0, 4.0), (lx, lz), audio)
        gates = (20.0 <= out[3] <= 20000.0) and out[0] > 0 and (out[4] != 0 or out[5] != 0)
        assert (np.abs(bufs[0]).sum() > 0) == gates
        if gates:
            n_ok += 1
            rt60 = out[2]
            assert (np.abs(bufs[1]).sum() > 0) == (rt60 <= 1.0)          # bus A: 0.5 s reverb, silent above T_ER_2
            assert (np.abs(bufs[3]).sum() > 0) == (rt60 >= 1.0)          # bus C: 3 s reverb, silent below T_ER_2
    assert n_ok > 50
    assert np.abs(pvdspref.render(np.zeros(8, np.float32), (3.0, 4.0), (lx, lz), audio)).sum() == 0
