"""Full-size parity (run on the B200 box: pytest -m gpu): the configurations bench.py and BASELINE.json quote, EXACTLY as they
are benchmarked (same scene, grid, listener list, step count and auto-selected step kernel), against the oracle
(oracle/pv_oracle.c, pinned bit-for-bit to the unmodified reference by tests/test_oracle.py).  The smaller parity tests of
tests/test_gpu_parity.py cover breadth; these close the gap between "the kernel is right at 250 x 250" and "the number in the
bench line was produced by a run whose outputs equal the reference's".  A few minutes of host time each (the oracle is a CPU
program); everything goes through the C-ABI."""
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


def test_config3_exactly_as_benchmarked_against_the_oracle(pv, scenes):
    """BASELINE.json configs[2] as bench.py runs it: BigRoom.pv on 1024 x 1024 cells, the bench's 4 batched listener positions,
    4000 steps, auto-selected step kernel.  For ALL FOUR sources: onset delays equal, obstruction / wet gain / RT60 / both
    direction vectors bit-exact, low-pass within 1 ulp, pressure planes at t = 3, 1000, 3999 bit-exact (FDTD.cpp:87-236,
    Analyzer.cpp:48-431)."""
    import bench
    cfg = bench.WORKLOAD
    assert (cfg["scene"], cfg["n"], cfg["T"], cfg["sources"]) == ("BigRoom", 1024, 4000, 4)
    size, scale, boxes = bench.scene_inputs(cfg)
    listeners = bench.bench_listeners(cfg["sources"], scale)
    gpu = pv.Scene(size, size, cfg["resolution"], T=cfg["T"], max_sources=cfg["sources"])
    assert (gpu.gx, gpu.gy, gpu.T) == (1024, 1024, 4000)
    for b in boxes:
        gpu.add_aabb(*b)
    res, dly = gpu.solve(listeners)
    ora = pvoracle.OracleSim(size, size, cfg["resolution"], T=cfg["T"], efree=float(gpu.efree))
    for b in boxes:
        ora.add_aabb(*b)
    for i, l in enumerate(listeners):
        ora.results[:] = 0
        ora.generate(l)
        ora.analyze(l)
        for t in (3, 1000, 3999):
            assert common.bit_equal(gpu.pressure(t, source=i), ora.hist[t].reshape(1025, 1025)).all(), f"source {i}: pressure plane t={t}"
        assert (ora.delay < 3e38).sum() > 100000
        assert_results(res[i], dly[i], ora.results, ora.delay, exclude=ora.clamped.astype(bool))
    gpu.close()


def test_config4_full_length_against_the_oracle(pv, scenes):
    """BASELINE.json configs[3] grid and scene at full length: HugeRoom.pv on 2048 x 2048 cells, 4000 steps, one source (a
    shard of the 8-source job), auto-selected step kernel.  The oracle's 67 GB pressure history does not fit a host, so it keeps
    the history of a 256-row band (oracle.pvoracle.OracleSim.generate_band): delays, obstruction, wet gain, low-pass and both
    direction vectors are checked for EVERY cell, RT60 and the pressure planes for the band."""
    size, scale = common.scaled_config(2048)
    boxes = common.boxes_of(scenes, "HugeRoom", scale)
    listener = common.listeners_for(1, scale)[0]
    gpu = pv.Scene(size, size, 275, T=4000, max_sources=1, efree=0.0447895788)
    ora = pvoracle.OracleSim(size, size, 275, T=4000, efree=0.0447895788)
    for b in boxes:
        gpu.add_aabb(*b); ora.add_aabb(*b)
    res, dly = gpu.solve([listener])
    row0, rows = 320, 256
    ora.generate_band(listener, row0, rows)
    ora.analyze(listener)
    for t in (3, 1000, 3999):
        assert common.bit_equal(gpu.pressure(t)[row0:row0 + rows], ora.hist[t].reshape(rows, 2049)).all(), f"pressure band t={t}"
    band = np.zeros((2048, 2048), bool)
    band[row0:row0 + rows] = True
    band = band.reshape(-1)
    valid = ora.delay < 3e38
    assert (valid & band).sum() > 100000
    excl = ora.clamped.astype(bool)
    # RT60 where the oracle computed it; everything else everywhere (a NaN RT60 outside the band is "not computed")
    got = np.array(res[0], copy=True)
    ref = np.array(ora.results, copy=True)
    got[~band, 2] = 0.0
    ref[~band, 2] = 0.0
    assert_results(got, dly[0], ref, ora.delay, exclude=excl)
    gpu.close()
