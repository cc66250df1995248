"""Multi-GPU form of the scene solver through the C-ABI (pvx_create_multi / pvx_multi_solve, include/planeverb_ext.h): listener
positions sharded contiguously over devices, one host thread per device, outputs gathered into one host table.  On a one-GPU
box the same code path is exercised by naming device 0 several times (an N-GPU run must equal N single-GPU runs
bit for bit); with more GPUs visible (gpurun --gpus N) every device takes part."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pv():
    from planeverb_b200 import pvcuda
    if pvcuda.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path, GPU tests cannot run")
    return pvcuda


def _reference_outputs(pv, size, boxes, listeners, emitters, T):
    """every listener solved alone on device 0, emitter outputs through pvx_lookup"""
    sc = pv.Scene(size, size, 275, T=T, max_sources=1)
    for b in boxes:
        sc.add_aabb(*b)
    out = np.full((len(listeners), len(emitters), 8), -1.0, np.float32)
    for i, l in enumerate(listeners):
        sc.clear_results(0)
        sc.solve([l], fetch=False)
        for e, pos in enumerate(emitters):
            v = sc.lookup(pos)
            if v is not None:
                out[i, e] = v
    sc.close()
    return out


@pytest.mark.parametrize("parts,max_batch", [(1, 0), (2, 0), (3, 1), (2, 2)])
def test_multi_solve_equals_single_solves(pv, scenes, parts, max_batch):
    n, T, S = 300, 240, 5
    size, scale = common.scaled_config(n)
    boxes = common.boxes_of(scenes, "FloorPlanScene", scale)
    listeners = common.listeners_for(S, scale)
    emitters = [(x * scale, 0.0, z * scale) for (x, z) in common.EMITTERS] + [(-1.0, 0.0, 2.0)]      # the last one is outside the grid
    ndev = pv.device_count()
    devices = [k % ndev for k in range(parts)]
    m = pv.MultiScene(devices, size, size, 275, T=T, max_sources=S, max_batch=max_batch, max_emitters=len(emitters))
    assert m.n_devices == parts
    for b in boxes:
        m.add_aabb(*b)
    got = m.solve(listeners, emitters)
    again = m.solve(listeners, emitters)                       # a second frame on the same scenes: same answer
    # every batch starts from zeroed result slots (a slot serves a different listener each time), like the first frame after Init
    want = _reference_outputs(pv, size, boxes, listeners, emitters, T)
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32))
    assert (got[:, -1] == -1.0).all()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # fewer listeners than devices: the empty shards are skipped
    one = m.solve(listeners[:1], emitters)
    assert np.array_equal(one.view(np.uint32), want[:1].view(np.uint32))
    m.close()


def test_multi_uses_every_visible_device(pv, scenes):
    """with N GPUs visible (gpurun --gpus N) the scenes really live on N different devices"""
    ndev = pv.device_count()
    size, scale = common.scaled_config(200)
    m = pv.MultiScene(list(range(ndev)), size, size, 275, T=120, max_sources=2 * ndev, max_emitters=2)
    assert m.n_devices == ndev and all(b == 2 for b in m.batches)
    # 2 listeners per device, all inside the 25 m (pre-scale) grid whatever the device count
    listeners = [((3.0 + 2.5 * (i % 8)) * scale, 0.0, (4.0 + 5.0 * (i // 8)) * scale) for i in range(2 * ndev)]
    emitters = [(5 * scale, 0, 6 * scale), (6 * scale, 0, 5 * scale)]
    got = m.solve(listeners, emitters)
    want = _reference_outputs(pv, size, [], listeners, emitters, 120)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    m.close()


def test_multi_error_paths(pv):
    with pytest.raises(pv.PlaneverbCudaError):
        pv.MultiScene([99], 25.0, 25.0, 275)
    m = pv.MultiScene([0], 25.0, 25.0, 275, max_sources=2, max_emitters=1)
    with pytest.raises(pv.PlaneverbCudaError):
        m.solve([(5, 0, 4)] * 3, [(5, 0, 6)])                 # more listeners than max_sources
    with pytest.raises(pv.PlaneverbCudaError) as e:
        m.solve([(500.0, 0, 4)], [(5, 0, 6)])                 # listener outside the grid: the device names itself
    assert "device 0" in str(e.value)
    assert m.solve([(5, 0, 4)], [(5, 0, 6)]).shape == (1, 1, 8)
    m.close()
