"""ctypes binding of planeverb_b200/lib/libplaneverb_b200.so -- the host-side mirror used by tests and
bench.py.  It binds the C-ABI declared in include/planeverb_cuda.h (pvc_*) and include/planeverb_ext.h
(pvx_*) and nothing else: no torch types cross the boundary, and there is NO CPU fallback -- if the
library is missing or no CUDA device is usable, every constructor raises.

`Scene` mirrors how a host program drives the reference's Grid / FreeGrid / Analyzer trio
(ProjectPlaneverb/src/FDTD/Grid.cpp, FreeGrid.cpp, src/DSP/Analyzer.cpp): add_aabb / remove_aabb,
generate+analyze for listener positions, per-emitter lookup.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PVC_LIB_PATH") or os.path.join(_HERE, "lib", "libplaneverb_b200.so")
CSRC = os.path.join(_HERE, "csrc")

PVC_OK, PVC_ERR_INVALID, PVC_ERR_MEMORY, PVC_ERR_CUDA, PVC_ERR_NO_DEVICE = range(5)
_STATUS = {1: "invalid argument/config", 2: "device memory", 3: "CUDA error", 4: "no CUDA device"}

# every symbol include/planeverb_cuda.h and include/planeverb_ext.h declare
PVC_SYMBOLS = [
    "pvc_device_count", "pvc_device_memory", "pvc_last_error", "pvc_create", "pvc_destroy", "pvc_memory_requirement",
    "pvc_create_streamed", "pvc_memory_requirement_streamed",
    "pvc_set_pulse", "pvc_clear_geometry", "pvc_apply_geometry", "pvc_fetch_coefficients",
    "pvc_compute_efree", "pvc_set_efree", "pvc_run", "pvc_synchronize", "pvc_clear_results",
    "pvc_fetch_results", "pvc_fetch_results_async", "pvc_fetch_wait", "pvc_fetch_result_at", "pvc_gather_results_async", "pvc_gather_wait", "pvc_fetch_ir", "pvc_fetch_pressure", "pvc_fetch_state",
    "pvc_last_timing", "pvc_last_launch_counts", "pvc_results_dev", "pvc_stream", "pvc_host_alloc", "pvc_host_free",
    "pvc_mark", "pvc_mark_elapsed", "pvc_debug_timeline", "pvc_debug_ws2_item", "pvc_debug_ws2_item_banded", "pvc_set_walk_mode", "pvc_step_variant",
]
PVX_SYMBOLS = [
    "pvx_create", "pvx_create_streamed", "pvx_history_steps", "pvx_destroy", "pvx_info", "pvx_pulse", "pvx_add_aabb", "pvx_remove_aabb",
    "pvx_flush_geometry", "pvx_solve", "pvx_solve_async", "pvx_wait", "pvx_solve_pipelined", "pvx_fetch_wait", "pvx_lookup", "pvx_lookup_async", "pvx_lookup_wait",
    "pvx_impulse_response", "pvx_solver",
    "pvx_create_multi", "pvx_destroy_multi", "pvx_multi_devices", "pvx_multi_scene", "pvx_multi_batch", "pvx_multi_add_aabb",
    "pvx_multi_remove_aabb", "pvx_multi_solve", "pvx_multi_last_error", "pvx_shard_bounds",
    "pvx_derive", "pvx_derive_pulse", "pvx_derive_rect", "pvx_derive_listener", "pvx_derive_emitter_cell",
]


class PvcConfig(C.Structure):
    _fields_ = [("gx", C.c_int), ("gy", C.c_int), ("T", C.c_int), ("fs", C.c_int), ("resolution", C.c_int),
                ("dx", C.c_float), ("courant", C.c_float), ("flux_samples", C.c_int), ("dry_samples", C.c_int),
                ("wet_samples", C.c_int), ("tail_samples", C.c_int), ("max_sources", C.c_int), ("device", C.c_int),
                ("step_kernel", C.c_int), ("reserved", C.c_int)]


class PvcRect(C.Structure):
    _fields_ = [("r0", C.c_int), ("r1", C.c_int), ("c0", C.c_int), ("c1", C.c_int), ("add", C.c_int),
                ("admittance", C.c_float)]


class PvcListener(C.Structure):
    _fields_ = [("cell_r", C.c_int), ("cell_c", C.c_int), ("efree_r", C.c_int), ("efree_c", C.c_int),
                ("x", C.c_float), ("z", C.c_float)]


class PlaneverbCudaError(RuntimeError):
    pass


def build(verbose=False):
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


_lib = None
_f, _i, _vp = C.c_float, C.c_int, C.c_void_p


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlaneverbCudaError(
                f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                "planeverb_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.pvc_last_error.restype = C.c_char_p
        L.pvc_memory_requirement.restype = C.c_size_t
        L.pvc_memory_requirement_streamed.restype = C.c_size_t
        L.pvc_memory_requirement_streamed.argtypes = [_vp, _i]
        L.pvc_results_dev.restype = _vp
        L.pvc_stream.restype = _vp
        L.pvx_solver.restype = _vp
        L.pvc_host_alloc.restype = _vp
        L.pvc_host_alloc.argtypes = [C.c_size_t]
        L.pvc_host_free.argtypes = [_vp]
        L.pvx_create.argtypes = [_f, _f, _i, _i, _f, _i, _i, _i, _i, _vp]
        L.pvx_create_streamed.argtypes = [_f, _f, _i, _i, _f, _i, _i, _i, _i, _i, _vp]
        L.pvx_destroy.argtypes = [_vp]
        L.pvx_history_steps.argtypes = [_vp]
        L.pvx_info.argtypes = [_vp, _vp, _vp]
        L.pvx_pulse.argtypes = [_vp, _vp, _i]
        L.pvx_add_aabb.argtypes = [_vp] + [_f] * 5
        L.pvx_remove_aabb.argtypes = [_vp] + [_f] * 5
        L.pvx_flush_geometry.argtypes = [_vp]
        L.pvx_solve.argtypes = [_vp, _vp, _i, _i, _vp, _vp]
        L.pvx_solve_async.argtypes = [_vp, _vp, _i, _i]
        L.pvx_wait.argtypes = [_vp]
        L.pvx_solve_pipelined.argtypes = [_vp, _vp, _i, _vp, _vp]
        L.pvx_fetch_wait.argtypes = [_vp]
        L.pvc_fetch_results_async.argtypes = [_vp, _i, _vp, _vp]
        L.pvc_fetch_wait.argtypes = [_vp]
        L.pvx_lookup.argtypes = [_vp, _i, _f, _f, _f, _vp]
        L.pvx_lookup_async.argtypes = [_vp, _i, _vp, _i, _vp, _vp]
        L.pvx_lookup_wait.argtypes = [_vp, _i]
        L.pvc_gather_results_async.argtypes = [_vp, _i, _vp, _i, _vp, _vp]
        L.pvc_gather_wait.argtypes = [_vp, _i]
        L.pvx_impulse_response.argtypes = [_vp, _i, _f, _f, _f, _vp]
        L.pvx_solver.argtypes = [_vp]
        L.pvc_fetch_results.argtypes = [_vp, _i, _vp, _vp]
        L.pvc_fetch_ir.argtypes = [_vp, _i, _i, _i, _vp]
        L.pvc_fetch_pressure.argtypes = [_vp, _i, _i, _vp]
        L.pvc_fetch_state.argtypes = [_vp, _i, _vp, _vp, _vp]
        L.pvc_fetch_coefficients.argtypes = [_vp, _vp, _vp]
        L.pvc_last_timing.argtypes = [_vp, _vp, _vp]
        if hasattr(L, "pvc_last_launch_counts"):          # absent from older builds loaded through PVC_LIB_PATH (A/B timing)
            L.pvc_last_launch_counts.argtypes = [_vp, _vp, _vp]
        L.pvc_clear_results.argtypes = [_vp, _i]
        L.pvc_synchronize.argtypes = [_vp]
        L.pvc_mark.argtypes = [_vp, _i]
        L.pvc_mark_elapsed.argtypes = [_vp, _vp]
        L.pvc_clear_geometry.argtypes = [_vp]
        L.pvc_set_walk_mode.argtypes = [_vp, _i]
        L.pvx_create_multi.argtypes = [_vp, _i, _f, _f, _i, _i, _f, _i, _i, _i, _vp]
        L.pvx_destroy_multi.argtypes = [_vp]
        L.pvx_multi_devices.argtypes = [_vp]
        L.pvx_multi_scene.argtypes = [_vp, _i]
        L.pvx_multi_scene.restype = _vp
        L.pvx_multi_batch.argtypes = [_vp, _i]
        L.pvx_multi_add_aabb.argtypes = [_vp] + [_f] * 5
        L.pvx_multi_remove_aabb.argtypes = [_vp] + [_f] * 5
        L.pvx_multi_solve.argtypes = [_vp, _vp, _i, _vp, _i, _vp]
        L.pvx_multi_last_error.argtypes = [_vp]
        L.pvx_multi_last_error.restype = C.c_char_p
        L.pvx_shard_bounds.argtypes = [_i, _i, _i, _vp, _vp]
        L.pvc_step_variant.argtypes = [_vp]
        _lib = L
    return _lib


def derive(resolution, size_x, size_y, T=0):
    """Host-side derivation (no GPU needed): returns (PvcConfig, dt, free_radius, free_ints[5])."""
    cfg = PvcConfig()
    ff = np.zeros(2, np.float32)
    ii = np.zeros(5, np.int32)
    _check(lib().pvx_derive(int(resolution), C.c_float(size_x), C.c_float(size_y), int(T), C.byref(cfg), _p(ff), _p(ii)), "pvx_derive")
    return cfg, np.float32(ff[0]), np.float32(ff[1]), [int(v) for v in ii]


def derive_pulse(resolution, fs, n):
    out = np.zeros(n, np.float32)
    _check(lib().pvx_derive_pulse(int(resolution), int(fs), _p(out), int(n)), "pvx_derive_pulse")
    return out


def derive_rect(resolution, px, py, w, h, absorption, add=True):
    r = PvcRect()
    _check(lib().pvx_derive_rect(int(resolution), C.c_float(px), C.c_float(py), C.c_float(w), C.c_float(h),
                                 C.c_float(absorption), int(add), C.byref(r)), "pvx_derive_rect")
    return r


def derive_listener(resolution, x, z):
    l = PvcListener()
    _check(lib().pvx_derive_listener(int(resolution), C.c_float(x), C.c_float(z), C.byref(l)), "pvx_derive_listener")
    return l


def derive_emitter_cell(resolution, size_x, size_y, x, z):
    rc = np.zeros(2, np.int32)
    code = lib().pvx_derive_emitter_cell(int(resolution), C.c_float(size_x), C.c_float(size_y), C.c_float(x), C.c_float(z), _p(rc))
    return None if code else (int(rc[0]), int(rc[1]))


def memory_requirement(gx, gy, T, max_sources, variant=0, step_kernel=0, history_steps=0):
    """Device bytes a solver of this size allocates (pvc_memory_requirement / pvc_memory_requirement_streamed: host arithmetic,
    needs no GPU); 0 = invalid config.  history_steps > 0: a streamed solver whose history holds that many samples."""
    cfg = PvcConfig(gx=int(gx), gy=int(gy), T=int(T), fs=1443, resolution=275, dx=0.3565818, courant=0.6666667, flux_samples=7, dry_samples=14,
                    wet_samples=115, tail_samples=14, max_sources=int(max_sources), device=0, step_kernel=int(step_kernel),
                    reserved=int(variant))
    if history_steps > 0:
        return int(lib().pvc_memory_requirement_streamed(C.byref(cfg), int(history_steps)))
    return int(lib().pvc_memory_requirement(C.byref(cfg)))


def device_count():
    return int(lib().pvc_device_count())


def device_memory(device=0):
    """(free, total) bytes of device memory (pvc_device_memory)"""
    f, t = C.c_size_t(), C.c_size_t()
    _check(lib().pvc_device_memory(int(device), C.byref(f), C.byref(t)), "pvc_device_memory")
    return int(f.value), int(t.value)


def _check(rc, what):
    if rc != 0:
        msg = lib().pvc_last_error().decode(errors="replace")
        raise PlaneverbCudaError(f"{what}: {_STATUS.get(rc, rc)}" + (f" ({msg})" if msg else ""))


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def pinned_array(shape, dtype=np.float32):
    """numpy array over page-locked host memory (freed when the array is collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    ptr = lib().pvc_host_alloc(n * dtype.itemsize)
    if not ptr:
        _check(PVC_ERR_MEMORY, "pvc_host_alloc")
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[id(buf)] = (buf, ptr)
    import weakref
    weakref.finalize(arr, _free_pinned, id(buf))
    return arr


_PINNED = {}


def _free_pinned(key):
    ent = _PINNED.pop(key, None)
    if ent is not None and _lib is not None:
        _lib.pvc_host_free(ent[1])


class Scene:
    """Grid + FreeGrid + Analyzer on one B200 (batched listeners)."""

    def __init__(self, size_x, size_y, resolution, T=0, efree=-1.0, max_sources=1, device=0,
                 step_kernel=0, variant=0, history_steps=0):
        """history_steps > 0: streamed solver (pvx_create_streamed) -- the pressure history holds that many samples and the
        response is solved in chunks; same results, for responses whose full history does not fit the device."""
        h = _vp()
        _check(lib().pvx_create_streamed(size_x, size_y, int(resolution), int(T), float(efree), int(max_sources),
                                         int(device), int(step_kernel), int(variant), int(history_steps), C.byref(h)), "pvx_create")
        self._h = h
        ii = np.zeros(10, np.int32)
        ff = np.zeros(4, np.float32)
        _check(lib().pvx_info(self._h, _p(ii), _p(ff)), "pvx_info")
        (self.gx, self.gy, self.T, self.fs, self.Sd, self.D, self.W, self.tail, self.free_samples,
         self.max_sources) = (int(v) for v in ii)
        self.dx, self.dt, self.courant, self.efree = (np.float32(v) for v in ff)
        self.resolution = int(resolution)
        self.size_x, self.size_y = float(size_x), float(size_y)
        self._solver = lib().pvx_solver(self._h)
        self.history_steps = int(lib().pvx_history_steps(self._h))      # 0: full history; > 0: streamed solver

    def close(self):
        if getattr(self, "_h", None):
            lib().pvx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def pulse(self):
        out = np.zeros(self.T, np.float32)
        _check(lib().pvx_pulse(self._h, _p(out), self.T), "pvx_pulse")
        return out

    def add_aabb(self, px, py, w, h, absorption):
        _check(lib().pvx_add_aabb(self._h, px, py, w, h, absorption), "pvx_add_aabb")

    def remove_aabb(self, px, py, w, h, absorption=0.0):
        _check(lib().pvx_remove_aabb(self._h, px, py, w, h, absorption), "pvx_remove_aabb")

    def flush_geometry(self):
        _check(lib().pvx_flush_geometry(self._h), "pvx_flush_geometry")

    def coef(self):
        self.flush_geometry()
        n = (self.gx + 1) * (self.gy + 1)
        b = np.zeros(n, np.int16)
        y = np.zeros(n, np.float32)
        _check(lib().pvc_fetch_coefficients(self._solver, _p(b), _p(y)), "pvc_fetch_coefficients")
        shp = (self.gx + 1, self.gy + 1)
        return b.reshape(shp), y.reshape(shp)

    @staticmethod
    def _listeners(listeners):
        a = np.ascontiguousarray(np.asarray(listeners, np.float32).reshape(-1, 3))
        return a, a.shape[0]

    def solve(self, listeners, analyze=True, fetch=True, out=None):
        """GenerateResponse + AnalyzeResponses for each listener (x, y, z). Returns (results, delay)
        with shapes (n, gx*gy, 8) and (n, gx*gy) when fetch, else None. out = (results, delay) host
        buffers to fill (e.g. pinned_array) instead of fresh ones."""
        a, n = self._listeners(listeners)
        cells = self.gx * self.gy
        if out is not None:
            res, dly = out
            assert res.shape == (n, cells, 8) and dly.shape == (n, cells)
            fetch = True
        else:
            res = np.zeros((n, cells, 8), np.float32) if fetch else None
            dly = np.zeros((n, cells), np.float32) if fetch else None
        _check(lib().pvx_solve(self._h, _p(a), n, int(bool(analyze)), _p(res), _p(dly)), "pvx_solve")
        return (res, dly) if fetch else None

    def solve_async(self, listeners, analyze=True):
        a, n = self._listeners(listeners)
        _check(lib().pvx_solve_async(self._h, _p(a), n, int(bool(analyze))), "pvx_solve_async")

    def wait(self):
        _check(lib().pvx_wait(self._h), "pvx_wait")

    def solve_pipelined(self, listeners, out):
        """Frame-loop form: enqueue the solve and the copy of ITS result grids into out = (results, delay) (pinned_array
        buffers) and return at once; the copy overlaps the next solve.  fetch_wait() returns when out is filled."""
        a, n = self._listeners(listeners)
        res, dly = out
        cells = self.gx * self.gy
        assert res.shape == (n, cells, 8) and dly.shape == (n, cells)
        _check(lib().pvx_solve_pipelined(self._h, _p(a), n, _p(res), _p(dly)), "pvx_solve_pipelined")

    def fetch_wait(self):
        _check(lib().pvx_fetch_wait(self._h), "pvx_fetch_wait")

    def emitter_cell(self, pos):
        """Analyzer::GetResponseResult's cell of a world-space emitter position (Analyzer.cpp:106-116), or None outside the grid."""
        return derive_emitter_cell(self.resolution, self.size_x, self.size_y, pos[0], pos[2])

    def fetch_results(self, source=0):
        cells = self.gx * self.gy
        res = np.zeros((cells, 8), np.float32)
        dly = np.zeros(cells, np.float32)
        _check(lib().pvc_fetch_results(self._solver, int(source), _p(res), _p(dly)), "pvc_fetch_results")
        return res, dly

    def clear_results(self, source=0):
        _check(lib().pvc_clear_results(self._solver, int(source)), "pvc_clear_results")

    def lookup(self, pos, source=0):
        out = np.zeros(8, np.float32)
        rc = lib().pvx_lookup(self._h, int(source), float(pos[0]), float(pos[1]), float(pos[2]), _p(out))
        if rc == PVC_ERR_INVALID:
            return None
        _check(rc, "pvx_lookup")
        return out

    def lookup_async(self, emitters, out, n=None):
        """Frame-loop form of lookup(): enqueue, in stream order after the last solve, the copy of the outputs of the emitter
        positions [(x, y, z), ...] for sources 0..n-1 into out (pinned_array of shape (n, len(emitters), 8)); returns a ticket
        for lookup_wait().  Emitters outside the grid read as eight -1."""
        n = int(n if n is not None else out.shape[0])
        e = np.ascontiguousarray(np.asarray(emitters, np.float32).reshape(-1, 3))
        assert out.shape == (n, e.shape[0], 8) and out.dtype == np.float32
        t = C.c_int(0)
        _check(lib().pvx_lookup_async(self._h, n, _p(e), int(e.shape[0]), _p(out), C.byref(t)), "pvx_lookup_async")
        return int(t.value)

    def lookup_wait(self, ticket):
        _check(lib().pvx_lookup_wait(self._h, int(ticket)), "pvx_lookup_wait")

    def impulse_response(self, pos, source=0):
        out = np.zeros((self.T, 3), np.float32)
        _check(lib().pvx_impulse_response(self._h, int(source), float(pos[0]), float(pos[1]), float(pos[2]), _p(out)),
               "pvx_impulse_response")
        return out

    def ir(self, r, c, source=0):
        out = np.zeros((self.T, 3), np.float32)
        _check(lib().pvc_fetch_ir(self._solver, int(source), int(r), int(c), _p(out)), "pvc_fetch_ir")
        return out

    def pressure(self, t, source=0):
        out = np.zeros((self.gx + 1, self.gy + 1), np.float32)
        _check(lib().pvc_fetch_pressure(self._solver, int(source), int(t), _p(out)), "pvc_fetch_pressure")
        return out

    def state(self, source=0):
        shp = (self.gx + 1, self.gy + 1)
        p, vx, vy = (np.zeros(shp, np.float32) for _ in range(3))
        _check(lib().pvc_fetch_state(self._solver, int(source), _p(p), _p(vx), _p(vy)), "pvc_fetch_state")
        return p, vx, vy

    def set_walk_mode(self, sequential):
        """listener direction by the reference's sequential walk (True) or pointer jumping (False, default)"""
        _check(lib().pvc_set_walk_mode(self._solver, int(bool(sequential))), "pvc_set_walk_mode")

    def step_variant(self):
        """the step-kernel variant this scene's solver runs (auto-selection resolved)"""
        return int(lib().pvc_step_variant(self._solver))

    def clear_geometry(self):
        _check(lib().pvc_clear_geometry(self._solver), "pvc_clear_geometry")

    def mark(self, which):
        _check(lib().pvc_mark(self._solver, int(which)), "pvc_mark")

    def mark_elapsed_ms(self):
        ms = C.c_float()
        _check(lib().pvc_mark_elapsed(self._solver, C.byref(ms)), "pvc_mark_elapsed")
        return float(ms.value)

    def timing(self):
        """(step_ms, analyzer_ms, total_ms, kernel_launches) of the last solve, CUDA events on the solver stream."""
        out = np.zeros(3, np.float32)
        n = C.c_int()
        _check(lib().pvc_last_timing(self._solver, _p(out), C.byref(n)), "pvc_last_timing")
        return float(out[0]), float(out[1]), float(out[2]), int(n.value)

    def launch_counts(self):
        """(step-phase kernel launches, analyzer-phase kernel launches) of the last solve."""
        a, b = C.c_int(), C.c_int()
        _check(lib().pvc_last_launch_counts(self._solver, C.byref(a), C.byref(b)), "pvc_last_launch_counts")
        return int(a.value), int(b.value)


def shard_bounds(n_items, parts, part):
    """the C-ABI's sharding rule (pvx_shard_bounds, host arithmetic): contiguous balanced shard [lo, hi)"""
    lo, hi = C.c_int(), C.c_int()
    _check(lib().pvx_shard_bounds(int(n_items), int(parts), int(part), C.byref(lo), C.byref(hi)), "pvx_shard_bounds")
    return lo.value, hi.value


class MultiScene:
    """pvx_create_multi / pvx_multi_solve: one scene per device, listeners sharded contiguously, one host thread per device,
    per-emitter outputs gathered into one host table (no torch, no collective)."""

    def __init__(self, devices, size_x, size_y, resolution, T=0, efree=-1.0, max_sources=1, max_batch=0, max_emitters=8):
        dev = np.ascontiguousarray(np.asarray(devices, np.int32))
        h = _vp()
        _check(lib().pvx_create_multi(_p(dev), int(dev.size), size_x, size_y, int(resolution), int(T), float(efree), int(max_sources),
                                      int(max_batch), int(max_emitters), C.byref(h)), "pvx_create_multi")
        self._h = h
        self.n_devices = int(lib().pvx_multi_devices(h))
        self.batches = [int(lib().pvx_multi_batch(h, k)) for k in range(self.n_devices)]

    def close(self):
        if getattr(self, "_h", None):
            lib().pvx_destroy_multi(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def add_aabb(self, px, py, w, h, absorption):
        _check(lib().pvx_multi_add_aabb(self._h, px, py, w, h, absorption), "pvx_multi_add_aabb")

    def remove_aabb(self, px, py, w, h, absorption=0.0):
        _check(lib().pvx_multi_remove_aabb(self._h, px, py, w, h, absorption), "pvx_multi_remove_aabb")

    def solve(self, listeners, emitters):
        """-> float32 [n_listeners, n_emitters, 8]"""
        a = np.ascontiguousarray(np.asarray(listeners, np.float32).reshape(-1, 3))
        e = np.ascontiguousarray(np.asarray(emitters, np.float32).reshape(-1, 3))
        out = np.zeros((a.shape[0], e.shape[0], 8), np.float32)
        rc = lib().pvx_multi_solve(self._h, _p(a), int(a.shape[0]), _p(e), int(e.shape[0]), _p(out))
        if rc:
            raise PlaneverbCudaError(f"pvx_multi_solve: {_STATUS.get(rc, rc)} ({lib().pvx_multi_last_error(self._h).decode(errors='replace')})")
        return out
