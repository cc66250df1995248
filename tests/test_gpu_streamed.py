"""Streamed solve (pvc_create_streamed / pvx_create_streamed; run on the B200 box: pytest -m gpu): a pressure history that holds
only `history_steps` samples, the response solved in chunks -- forward sweep with the causal analyzer sums carried per cell and
the state checkpointed at every chunk start, then the chunks recomputed from their checkpoints in reverse order for the
backward Schroeder pass (FDTD.cpp:87-236, Analyzer.cpp:139-328).  It must give the full-history solver's outputs BIT FOR BIT
(and therefore the oracle's) for any chunk length; its purpose is capacity: BASELINE.json configs[3]'s eight 2048 x 2048
sources as ONE batch on one GPU, or responses longer than the device could record.  Everything goes through the C-ABI."""
import numpy as np
import pytest

from oracle import pvoracle
from tests import common
from tests.test_gpu_parity import assert_results

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pv():
    from planeverb_b200 import pvcuda
    if pvcuda.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path, GPU tests cannot run")
    return pvcuda


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32), np.ascontiguousarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("scene,n,T,history,nsrc,variant", [
    ("SmallRoom", None, 0, 104, 1, 0),        # Sandbox default 70 x 70, T = 435: K = 5 chunks, the last one 19 samples (resident 4-warp tiling)
    ("SmallRoom", None, 0, 104, 1, 47),       # ... on the generational kernel
    ("SmallRoom", None, 0, 8, 1, 0),          # the shortest chunk there is: K = 55, every analysis window straddles chunk borders
    ("SmallRoom", None, 0, 8, 1, 50),
    ("SmallRoom", None, 0, 1000, 1, 0),       # history longer than the response: one chunk, the chunk kernels against the full ones
    ("FloorPlanScene", 250, 600, 304, 3, 0),  # K = 2 (no checkpoint at all), three batched listeners, walls in most tiles
    ("FloorPlanScene", 250, 601, 200, 3, 47), # K = 4 with a one-sample last chunk
    ("FloorPlanScene", 250, 601, 200, 3, 65), # ... on 18-warp resident tiles (96 registers, the config-3 kernel)
    ("FloorPlanScene", 240, 601, 200, 2, 69), # ... on two-CTA 8-warp tiles, gx a multiple of the owned rows (live padding row across chunks)
    ("BigRoom", 512, 1203, 400, 2, 0),        # K = 4, T not a multiple of 4 (resident 12-warp tiling)
    ("BigRoom", 512, 1203, 400, 2, 47),
])
def test_streamed_solve_equals_the_full_history_solve_and_the_oracle(pv, scenes, scene, n, T, history, nsrc, variant):
    if n is None:
        size, scale = 25.0, 1.0
    else:
        size, scale = common.scaled_config(n)
    boxes = common.boxes_of(scenes, scene, scale)
    listeners = common.listeners_for(nsrc, scale)
    full = pv.Scene(size, size, 275, T=T, max_sources=nsrc)
    strm = pv.Scene(size, size, 275, T=T, max_sources=nsrc, history_steps=history, efree=float(full.efree), variant=variant)
    assert (strm.gx, strm.gy, strm.T) == (full.gx, full.gy, full.T) and strm.step_variant() == (variant or full.step_variant())
    for b in boxes:
        full.add_aabb(*b); strm.add_aabb(*b)
    rf, df = full.solve(listeners)
    rs, ds = strm.solve(listeners)
    assert np.array_equal(df, ds), "onset delays differ"
    assert (df < 3e38).sum() > 100
    for k, name in enumerate(common.FIELDS):
        assert same_bits(rf[:, :, k], rs[:, :, k]), f"{name}: {int((rf[:, :, k].view(np.uint32) != rs[:, :, k].view(np.uint32)).sum())} cells differ from the full-history solve"
    # a second solve on the same solver (other listeners, stale carry and checkpoints behind it) is just as exact
    moved = [(l[0] * 0.9, 0.0, l[2] * 1.1) for l in listeners]
    for i in range(nsrc):
        full.clear_results(i); strm.clear_results(i)
    rf2, df2 = full.solve(moved)
    rs2, ds2 = strm.solve(moved)
    assert np.array_equal(df2, ds2) and same_bits(rf2, rs2)
    # ... and the oracle agrees (source 0)
    ora = pvoracle.OracleSim(size, size, 275, T=T, efree=float(full.efree))
    for b in boxes:
        ora.add_aabb(*b)
    ora.generate(listeners[0]); ora.analyze(listeners[0])
    assert_results(rs[0], ds[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    full.close(); strm.close()


@pytest.mark.parametrize("n,T,history,listener_cell", [
    (130, 90, 32, (130, 40)),       # listener on the padding row: a dead source, its last-sample patch belongs to the LAST chunk only
    (119, 133, 40, (0, 0)),         # corner listener
    (97, 64, 24, (50, 50)),
])
def test_streamed_final_state_and_dead_sources(pv, n, T, history, listener_cell):
    size, _ = common.scaled_config(n)
    ora = pvoracle.OracleSim(size, size, 275, T=T, efree=0.0447895788)
    full = pv.Scene(size, size, 275, T=T, efree=0.0447895788)
    strm = pv.Scene(size, size, 275, T=T, efree=0.0447895788, history_steps=history)
    dx = float(ora.dx)
    L = ((listener_cell[0] + 0.5) * dx, 0.0, (listener_cell[1] + 0.5) * dx)
    for b in [(-0.5 * dx, 0.3 * n * dx, 3 * dx, 4 * dx, 0.9), (n * dx, 0.7 * n * dx, 4 * dx, 6 * dx, 0.5),
              (0.6 * n * dx, n * dx, 5 * dx, 2.5 * dx, 0.97)]:
        ora.add_aabb(*b); full.add_aabb(*b); strm.add_aabb(*b)
    # forward sweep only: the state planes hold the end of the response
    full.solve([L], analyze=False, fetch=False); full.wait()
    strm.solve([L], analyze=False, fetch=False); strm.wait()
    for a, b in zip(full.state(), strm.state()):
        assert common.bit_equal(a, b).all()          # the full-history solver may run another step kernel: exact zeros can differ in sign
    # analyzed: same outputs as the full-history solver and the oracle; the state planes are then mid-response and say so
    rf, df = full.solve([L])
    rs, ds = strm.solve([L])
    assert np.array_equal(df, ds) and same_bits(rf, rs)
    ora.generate(L); ora.analyze(L)
    assert_results(rs[0], ds[0], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    if T > history:
        with pytest.raises(pv.PlaneverbCudaError):
            strm.state()
    with pytest.raises(pv.PlaneverbCudaError):
        strm.pressure(3)
    with pytest.raises(pv.PlaneverbCudaError):
        strm.ir(5, 5)
    full.close(); strm.close()


def test_streamed_solver_rejects_what_it_cannot_run(pv):
    with pytest.raises(pv.PlaneverbCudaError):
        pv.Scene(25.0, 25.0, 275, history_steps=104, step_kernel=1)          # two-launch baseline: no chunked form
    with pytest.raises(pv.PlaneverbCudaError):
        pv.Scene(25.0, 25.0, 275, history_steps=104, variant=18)             # the plain one-launch-per-4-steps fallback: no chunked form
    # the free-field probe must fit the history
    with pytest.raises(pv.PlaneverbCudaError):
        pv.Scene(25.0, 25.0, 275, history_steps=8)
    s = pv.Scene(25.0, 25.0, 275, history_steps=8, efree=0.0447895788)
    s.close()


def test_config4_eight_sources_as_one_streamed_batch(pv, scenes):
    """BASELINE.json configs[3]: HugeRoom.pv on 2048 x 2048 cells, 8 sources, 4000 steps.  The full history is 71 GB per source
    -- a B200 holds two -- so the full-history solver runs the job as four batches; the streamed solver holds all eight with an
    800-sample history (K = 5 chunks).  Outputs must be bit-identical."""
    size, scale = common.scaled_config(2048)
    boxes = common.boxes_of(scenes, "HugeRoom", scale)
    listeners = common.listeners_for(8, scale)
    need = pv.memory_requirement(2048, 2048, 4000, 8, history_steps=800)
    free, total = pv.device_memory(0)
    if need > 0.95 * free:
        pytest.skip(f"streamed config 4 needs {need / 1e9:.0f} GB, device has {free / 1e9:.0f} GB free")
    strm = pv.Scene(size, size, 275, T=4000, max_sources=8, history_steps=800)
    efree = float(strm.efree)
    for b in boxes:
        strm.add_aabb(*b)
    rs, ds = strm.solve(listeners)
    st, an, tot, launches = strm.timing()
    strm.close()
    full = pv.Scene(size, size, 275, T=4000, max_sources=2, efree=efree)
    for b in boxes:
        full.add_aabb(*b)
    ms_full = 0.0
    for i in range(0, 8, 2):
        full.clear_results(0); full.clear_results(1)
        rf, df = full.solve(listeners[i:i + 2])
        ms_full += full.timing()[2]
        assert np.array_equal(df, ds[i:i + 2]), f"sources {i}, {i + 1}: onset delays differ"
        assert same_bits(rf, rs[i:i + 2]), f"sources {i}, {i + 1}: outputs differ"
        assert (df < 3e38).sum() > 1000000
    full.close()
    print(f"config 4 on one GPU: streamed batch of 8 (history 800) {tot:.1f} ms, {launches} launches; four full-history batches of 2: {ms_full:.1f} ms")


def test_automatic_history_length_falls_back_to_a_streamed_solver(pv, scenes):
    """history_steps = -1 (what Planeverb::Init and pvx_create_multi pass): the full history when it fits 90 % of the device's free
    memory, else the longest history that does.  Three 2048 x 2048 x 4000 sources need 214 GB of history: a streamed solver."""
    size, scale = common.scaled_config(2048)
    boxes = common.boxes_of(scenes, "HugeRoom", scale)
    listeners = common.listeners_for(3, scale)
    small = pv.Scene(25.0, 25.0, 275, history_steps=-1)
    assert small.history_steps == 0                          # fits: the ordinary solver
    small.close()
    free, _ = pv.device_memory(0)
    if pv.memory_requirement(2048, 2048, 4000, 3) <= 0.9 * free:
        pytest.skip("three full 2048^2 histories fit this device")
    auto = pv.Scene(size, size, 275, T=4000, max_sources=3, history_steps=-1)
    assert 0 < auto.history_steps < 4000 and auto.history_steps % 8 == 0
    assert pv.memory_requirement(2048, 2048, 4000, 3, history_steps=auto.history_steps) <= 0.9 * free
    for b in boxes:
        auto.add_aabb(*b)
    ra, da = auto.solve(listeners)
    efree = float(auto.efree)
    auto.close()
    full = pv.Scene(size, size, 275, T=4000, max_sources=1, efree=efree)
    for b in boxes:
        full.add_aabb(*b)
    rf, df = full.solve(listeners[2:3])
    assert np.array_equal(df[0], da[2]) and same_bits(rf[0], ra[2])
    full.close()


@pytest.mark.parametrize("scene,n,T,history,nsrc", [
    ("BigRoom", 1024, 16000, 2000, 1),      # a response four times config 3's length: K = 8
    ("HugeRoom", 4096, 1000, 248, 1),       # a grid four times config 4's area: K = 5 (its full 4000-step history would be 268 GB)
])
def test_streamed_long_responses_and_huge_grids(pv, scenes, scene, n, T, history, nsrc):
    """Sizes the streamed solver exists for, chosen so that the full history still fits the device once and the two solvers can be
    compared bit for bit."""
    size, scale = common.scaled_config(n)
    boxes = common.boxes_of(scenes, scene, scale)
    listeners = common.listeners_for(nsrc, scale)
    free, _ = pv.device_memory(0)
    need = pv.memory_requirement(n, n, T, nsrc) + pv.memory_requirement(n, n, T, nsrc, history_steps=history)
    if need > 0.95 * free:
        pytest.skip(f"needs {need / 1e9:.0f} GB of device memory, {free / 1e9:.0f} GB free")
    full = pv.Scene(size, size, 275, T=T, max_sources=nsrc, efree=0.0447895788)
    strm = pv.Scene(size, size, 275, T=T, max_sources=nsrc, efree=0.0447895788, history_steps=history)
    for b in boxes:
        full.add_aabb(*b); strm.add_aabb(*b)
    rf, df = full.solve(listeners)
    rs, ds = strm.solve(listeners)
    assert (df < 3e38).sum() > 100000
    assert np.array_equal(df, ds) and same_bits(rf, rs)
    print(f"{scene} {n}^2 x{nsrc} T={T}: full history {full.timing()[2]:.1f} ms (variant {full.step_variant()}), streamed (history {history}) {strm.timing()[2]:.1f} ms")
    full.close(); strm.close()
