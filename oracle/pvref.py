"""ctypes view of oracle/_ref/libpvref.so -- the UNMODIFIED reference Grid/FreeGrid/Analyzer
(ProjectPlaneverb/src/FDTD, src/DSP) compiled in place by oracle/refdriver/Makefile.

TEST INFRASTRUCTURE ONLY: imported by tests/, tools/make_golden.py and bench.py's reference arm /
cpu_baseline leg.  Never imported by planeverb_b200/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PVREF_LIB selects another build of the same unmodified sources (oracle/_ref/libpvref_fast.so: the shipped Release flags,
# timing only -- bench.py's second CPU row); the oracle proper is the strict build
LIB_PATH = os.environ.get("PVREF_LIB") or os.path.join(_HERE, "_ref", "libpvref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.pvref_create.restype = C.c_void_p
        L.pvref_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.c_float]
        L.pvref_destroy.argtypes = [C.c_void_p]
        L.pvref_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for n in ("pvref_add_aabb", "pvref_remove_aabb"):
            getattr(L, n).argtypes = [C.c_void_p] + [C.c_float] * 5
        for n in ("pvref_generate", "pvref_analyze"):
            getattr(L, n).argtypes = [C.c_void_p] + [C.c_float] * 3
        L.pvref_clear_results.argtypes = [C.c_void_p]
        L.pvref_copy_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pvref_lookup.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.pvref_lookup.restype = C.c_int
        L.pvref_copy_ir.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pvref_copy_snapshot.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pvref_copy_coef.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pvref_copy_pulse.argtypes = [C.c_void_p, C.c_void_p]
        L.pvref_efree_per_r.argtypes = [C.c_void_p] + [C.c_int] * 4
        L.pvref_efree_per_r.restype = C.c_float
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefSim:
    """One reference Grid + FreeGrid + Analyzer (no Context thread)."""

    def __init__(self, size_x, size_y, resolution, T=0, efree=-1.0):
        self._h = lib().pvref_create(size_x, size_y, resolution, int(T), float(efree))
        ii = np.zeros(8, np.int32)
        ff = np.zeros(4, np.float32)
        lib().pvref_info(self._h, _p(ii), _p(ff))
        self.gx, self.gy, self.T, self.fs, self.pulse_mismatch = (int(v) for v in ii[:5])
        self.dx, self.dt, self.efree, self.courant = (np.float32(v) for v in ff)

    def close(self):
        if self._h:
            lib().pvref_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def add_aabb(self, px, py, w, h, absorption):
        lib().pvref_add_aabb(self._h, px, py, w, h, absorption)

    def remove_aabb(self, px, py, w, h, absorption):
        lib().pvref_remove_aabb(self._h, px, py, w, h, absorption)

    def generate(self, listener):
        lib().pvref_generate(self._h, *[float(v) for v in listener])

    def analyze(self, listener):
        lib().pvref_analyze(self._h, *[float(v) for v in listener])

    def clear_results(self):
        lib().pvref_clear_results(self._h)

    def results(self):
        n = self.gx * self.gy
        res = np.zeros((n, 8), np.float32)
        delay = np.zeros(n, np.float32)
        lib().pvref_copy_results(self._h, _p(res), _p(delay))
        return res, delay

    def lookup(self, pos):
        out = np.zeros(8, np.float32)
        ok = lib().pvref_lookup(self._h, *[float(v) for v in pos], _p(out))
        return out if ok else None

    def ir(self, r, c):
        out = np.zeros((self.T, 3), np.float32)
        lib().pvref_copy_ir(self._h, int(r), int(c), _p(out))
        return out

    def snapshot(self, t):
        N = (self.gx + 1) * (self.gy + 1)
        p, vx, vy = (np.zeros(N, np.float32) for _ in range(3))
        lib().pvref_copy_snapshot(self._h, int(t), _p(p), _p(vx), _p(vy))
        shp = (self.gx + 1, self.gy + 1)
        return p.reshape(shp), vx.reshape(shp), vy.reshape(shp)

    def coef(self):
        N = (self.gx + 1) * (self.gy + 1)
        b = np.zeros(N, np.int16)
        a = np.zeros(N, np.float32)
        lib().pvref_copy_coef(self._h, _p(b), _p(a))
        shp = (self.gx + 1, self.gy + 1)
        return b.reshape(shp), a.reshape(shp)

    def pulse(self):
        out = np.zeros(self.T, np.float32)
        lib().pvref_copy_pulse(self._h, _p(out))
        return out

    def efree_per_r(self, lx, ly, ex, ey):
        return np.float32(lib().pvref_efree_per_r(self._h, lx, ly, ex, ey))
