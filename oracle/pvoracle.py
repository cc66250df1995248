"""ctypes view of oracle/_build/libpvoracle.so (oracle/pv_oracle.c, the plain-C restatement of the
reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by planeverb_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libpvoracle.so")

_lib = None
_f, _i, _vp = C.c_float, C.c_int, C.c_void_p


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libpvoracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.pvo_grid_params.argtypes = [_i, _vp, _vp, _vp]
        L.pvo_derived.argtypes = [_i, _f, _f, _vp]
        L.pvo_derived.restype = _f
        L.pvo_gaussian_pulse.argtypes = [_i, _f, _vp, C.c_uint]
        L.pvo_coef_init.argtypes = [_i, _i, _vp, _vp]
        L.pvo_add_aabb.argtypes = [_i, _i, _f, _vp, _vp] + [_f] * 5
        L.pvo_remove_aabb.argtypes = [_i, _i, _f, _vp, _vp] + [_f] * 4
        L.pvo_listener_cell.argtypes = [_f, _f, _f, _vp, _vp]
        L.pvo_simulate.argtypes = [_i, _i, _vp, _vp, _f, _i, _vp, _i] + [_vp] * 11 + [_i] * 3
        L.pvo_simulate_band.argtypes = [_i, _i, _vp, _vp, _f, _i, _vp, _i] + [_vp] * 4 + [C.c_longlong] * 2 + [_vp] * 5 + [_i] * 3
        L.pvo_encode_band.argtypes = [_i, _i, _i, _i, _f, _f, _f, _f, _vp, C.c_longlong, C.c_longlong] + [_vp] * 8
        L.pvo_efree.argtypes = [_i, _i, _i]
        L.pvo_efree.restype = _f
        L.pvo_efree_per_r.argtypes = [_f, _f, _i, _i, _i, _i]
        L.pvo_efree_per_r.restype = _f
        L.pvo_encode.argtypes = [_i, _i, _i, _i, _f, _f, _f, _f] + [_vp] * 9
        L.pvo_directions.argtypes = [_i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp]
        L.pvo_windows.argtypes = [_i, _vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def grid_params(resolution):
    dx, dt, fs = C.c_float(), C.c_float(), C.c_uint()
    lib().pvo_grid_params(resolution, C.byref(dx), C.byref(dt), C.byref(fs))
    return np.float32(dx.value), np.float32(dt.value), int(fs.value)


def derived(resolution, size_x, size_y):
    out = np.zeros(3, np.int32)
    courant = lib().pvo_derived(resolution, size_x, size_y, _p(out))
    return int(out[0]), int(out[1]), int(out[2]), np.float32(courant)


def gaussian_pulse(resolution, fs, n):
    out = np.zeros(n, np.float32)
    lib().pvo_gaussian_pulse(resolution, float(fs), _p(out), n)
    return out


def windows(fs):
    out = np.zeros(4, np.int32)
    lib().pvo_windows(fs, _p(out))
    return tuple(int(v) for v in out)   # Sd, D, W, tail


def size_for_cells(resolution, n):
    """gridSizeInMeters that the reference truncates to exactly n cells."""
    dx, _, _ = grid_params(resolution)
    size = np.float32((n + 0.5) * float(dx))
    g = derived(resolution, float(size), float(size))
    assert g[0] == n and g[1] == n, (g, n)
    return float(size)


class OracleSim:
    """Grid + FreeGrid + Analyzer restated (one listener at a time)."""

    def __init__(self, size_x, size_y, resolution, T=0, efree=-1.0):
        self.resolution = int(resolution)
        self.dx, self.dt, self.fs = grid_params(resolution)
        self.gx, self.gy, natT, self.courant = derived(resolution, size_x, size_y)
        self.T = int(T) if T and T > 0 else natT
        self.S = self.gy + 1
        self.N = (self.gx + 1) * self.S
        self.b = np.zeros(self.N, np.int16)
        self.R = np.zeros(self.N, np.float32)
        lib().pvo_coef_init(self.gx, self.gy, _p(self.b), _p(self.R))
        self.pulse = gaussian_pulse(resolution, self.fs, self.T)
        self.efree = np.float32(efree) if efree >= 0 else np.float32(lib().pvo_efree(resolution, self.gx, self.gy))
        self.Sd, self.D, self.W, self.tail = windows(self.fs)
        self.results = np.zeros((self.gx * self.gy, 8), np.float32)
        self.delay = np.zeros(self.gx * self.gy, np.float32)
        self.clamped = np.zeros(self.gx * self.gy, np.uint8)
        self.hist = None

    def add_aabb(self, px, py, w, h, absorption):
        lib().pvo_add_aabb(self.gx, self.gy, self.dx, _p(self.b), _p(self.R), px, py, w, h, absorption)

    def remove_aabb(self, px, py, w, h, absorption=0.0):
        lib().pvo_remove_aabb(self.gx, self.gy, self.dx, _p(self.b), _p(self.R), px, py, w, h)

    def listener_cell(self, listener):
        lr, lc = C.c_int(), C.c_int()
        lib().pvo_listener_cell(self.dx, float(listener[0]), float(listener[2]), C.byref(lr), C.byref(lc))
        return lr.value, lc.value

    def generate_band(self, listener, row0, rows):
        """generate() keeping the pressure history of alloc rows [row0, row0 + rows) only (self.hist: (T, rows * (gy + 1))): for
        grids whose full history does not fit the host.  analyze() then yields RT60 for the interior cells of those rows and
        NaN elsewhere; every other output is complete."""
        N, T = self.N, self.T
        self.p = np.zeros(N, np.float32)
        self.vx = np.zeros(N, np.float32)
        self.vy = np.zeros(N, np.float32)
        self.band = (int(row0) * self.S, int(rows) * self.S)
        self.hist = np.zeros((T, self.band[1]), np.float32)
        self.hvx = self.hvy = None
        self.onset = np.zeros(N, np.int32)
        self.edry, self.fx, self.fy, self.wet = (np.zeros(N, np.float32) for _ in range(4))
        lr, lc = self.listener_cell(listener)
        lib().pvo_simulate_band(self.gx, self.gy, _p(self.b), _p(self.R), self.courant, lr * self.S + lc,
                                _p(self.pulse), T, _p(self.p), _p(self.vx), _p(self.vy),
                                _p(self.hist), self.band[0], self.band[1],
                                _p(self.onset), _p(self.edry), _p(self.fx), _p(self.fy), _p(self.wet),
                                self.Sd, self.D, self.W)

    def generate(self, listener, keep_velocity=False):
        self.band = None
        N, T = self.N, self.T
        self.p = np.zeros(N, np.float32)
        self.vx = np.zeros(N, np.float32)
        self.vy = np.zeros(N, np.float32)
        self.hist = np.zeros((T, N), np.float32)
        self.hvx = np.zeros((T, N), np.float32) if keep_velocity else None
        self.hvy = np.zeros((T, N), np.float32) if keep_velocity else None
        self.onset = np.zeros(N, np.int32)
        self.edry, self.fx, self.fy, self.wet = (np.zeros(N, np.float32) for _ in range(4))
        lr, lc = self.listener_cell(listener)
        lib().pvo_simulate(self.gx, self.gy, _p(self.b), _p(self.R), self.courant, lr * self.S + lc,
                           _p(self.pulse), T, _p(self.p), _p(self.vx), _p(self.vy),
                           _p(self.hist), _p(self.hvx), _p(self.hvy),
                           _p(self.onset), _p(self.edry), _p(self.fx), _p(self.fy), _p(self.wet),
                           self.Sd, self.D, self.W)

    def analyze(self, listener):
        lx, lz = float(listener[0]), float(listener[2])
        if getattr(self, "band", None):
            lib().pvo_encode_band(self.gx, self.gy, self.T, self.fs, self.dx, self.efree, lx, lz,
                                  _p(self.hist), self.band[0], self.band[1], _p(self.onset), _p(self.edry), _p(self.fx), _p(self.fy),
                                  _p(self.wet), _p(self.results), _p(self.delay), _p(self.clamped))
        else:
            lib().pvo_encode(self.gx, self.gy, self.T, self.fs, self.dx, self.efree, lx, lz,
                             _p(self.hist), _p(self.onset), _p(self.edry), _p(self.fx), _p(self.fy), _p(self.wet),
                             _p(self.results), _p(self.delay), _p(self.clamped))
        lib().pvo_directions(self.gx, self.gy, self.T, self.fs, self.resolution, self.dx, lx, lz,
                             _p(self.results), _p(self.delay))

    def snapshot(self, t):
        shp = (self.gx + 1, self.gy + 1)
        return (self.hist[t].reshape(shp),
                self.hvx[t].reshape(shp) if self.hvx is not None else None,
                self.hvy[t].reshape(shp) if self.hvy is not None else None)

    def coef(self):
        shp = (self.gx + 1, self.gy + 1)
        return self.b.reshape(shp), self.R.reshape(shp)
