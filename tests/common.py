"""Shared helpers for the parity tests: scene fixtures, BASELINE-style scaled configs and comparisons.

The oracle (oracle/) is used here only as the checker."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES_JSON = os.path.join(ROOT, "planeverb_b200", "scenes", "scenes.json")
GOLDEN = os.path.join(ROOT, "tests", "golden")

FIELDS = ["occlusion", "wetGain", "rt60", "lowpass", "dirX", "dirY", "srcDirX", "srcDirY"]
DEFAULT_LISTENER = (5.0, 0.0, 4.0)          # PlaneverbSandbox/src/Editor/Editor.cpp:36
EMITTERS = [(5, 6), (6, 5), (3.5, 3.5), (12.5, 12.5), (20, 20)]
RTOL = 1e-4                                  # BASELINE.json north_star tolerance


from planeverb_b200.scenes import load_scenes, boxes_of          # noqa: E402,F401  (the product's own scene helpers)


def scaled_config(n, resolution=275):
    """(size_m, scale) so that the reference truncates to exactly n x n cells and the 25 m scenes fill it: the ORACLE's
    derivation (tests/test_abi.py checks that the product's planeverb_b200.scenes.scaled_config agrees)."""
    from oracle import pvoracle
    size = pvoracle.size_for_cells(resolution, n)
    dx, _, _ = pvoracle.grid_params(resolution)
    scale = n * float(dx) / 25.0
    return size, scale


def listeners_for(k, scale=1.0):
    """k deterministic listener positions: Sandbox default plus (+1.5i, 0, +0.75i) m pre-scale."""
    return [((5.0 + 1.5 * i) * scale, 0.0, (4.0 + 0.75 * i) * scale) for i in range(k)]


def bit_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (a == b)     # a == b lets +0 match -0


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        d = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    d[(a == b) | (np.isnan(a) & np.isnan(b))] = 0.0
    return d


def compare_results(got, got_delay, ref, ref_delay, exclude=None, rtol=RTOL, direction_rtol=None):
    """Assert the analyzer outputs agree: delays exactly, every output field of every cell that has an
    onset within rtol (relative), direction for every cell. Returns per-field max relative error."""
    assert np.array_equal(got_delay, ref_delay), "onset delays differ"
    valid = ref_delay < 3e38
    if exclude is not None:
        valid = valid & ~exclude
    report = {}
    for k, name in enumerate(FIELDS):
        m = np.ones_like(valid) if k in (4, 5) else valid
        if exclude is not None and k in (4, 5):
            m = ~exclude
        e = rel_err(got[m, k], ref[m, k])
        # components of unit vectors: compare absolutely against the unit length
        if k >= 4:
            e = np.abs(got[m, k].astype(np.float64) - ref[m, k].astype(np.float64))
        tol = rtol if (k not in (4, 5) or direction_rtol is None) else direction_rtol
        report[name] = float(e.max()) if e.size else 0.0
        assert report[name] <= tol, f"{name}: max error {report[name]:.3e} > {tol:.1e} ({int((e > tol).sum())} cells)"
    return report


GOLDEN_CASES = ["smallroom_70", "bigroom_70", "shoebox_70", "hugeroom_70", "floorplan_70",
                "singlewall_95_res375", "middlewall_127_res500", "smallroom_128_T500",
                "directiontester_101_T300"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def golden_boxes(z):
    return [tuple(float(v) for v in row) for row in z["boxes"]]


def reference_clamped(meta, delay, D):
    """cells whose dry window runs past the IR (the reference reads out of bounds there)"""
    return (delay < 3e38) & (delay + D >= meta["T"])
