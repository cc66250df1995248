"""ctypes view of oracle/_ref/libpvdspref.so -- the UNMODIFIED reference consumer of the hot path's outputs,
PlaneverbDSP (PlaneverbDSP/src/PvDSPContext.cpp, DSP/Lowpass.cpp), compiled in place by oracle/dspdriver/Makefile.

TEST INFRASTRUCTURE ONLY: imported by tests/.  Never imported by planeverb_b200/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpvdspref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.pvdsp_render.restype = C.c_int
        L.pvdsp_render.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint, C.c_void_p, C.c_uint, C.c_int] + [C.c_void_p] * 4
        _lib = L
    return _lib


def test_signal(num_frames, seed=7):
    """deterministic stereo test signal, interleaved: a decaying chirp plus a little noise"""
    rng = np.random.RandomState(seed)
    t = np.arange(num_frames, dtype=np.float64) / 44100.0
    mono = np.sin(2 * np.pi * (200.0 + 8000.0 * t) * t) * np.exp(-3.0 * t) + 0.05 * rng.standard_normal(num_frames)
    st = np.stack([mono, 0.8 * mono], axis=1).astype(np.float32)
    return np.ascontiguousarray(st.reshape(-1))


def render(out8, emitter_xz, listener_xz, audio, sampling_rate=44100, calls=3):
    """Context::SubmitSource (PvDSPContext.cpp:250-425) for one emitter with acoustic parameters out8 (PlaneverbOutput order),
    `calls` audio callbacks.  Returns the (dry, A, B, C) stereo buffers of the last callback, shape (4, 2*frames).
    An input the reference's gates reject (:258-262) leaves all four buffers zero."""
    out8 = np.ascontiguousarray(out8, np.float32)
    audio = np.ascontiguousarray(audio, np.float32)
    frames = audio.size // 2
    bufs = np.zeros((4, 2 * frames), np.float32)
    ptr = [bufs[i].ctypes.data_as(C.c_void_p) for i in range(4)]
    rc = lib().pvdsp_render(out8.ctypes.data_as(C.c_void_p), float(emitter_xz[0]), float(emitter_xz[1]),
                            float(listener_xz[0]), float(listener_xz[1]), int(sampling_rate),
                            audio.ctypes.data_as(C.c_void_p), int(frames), int(calls), *ptr)
    if rc != 0:
        raise RuntimeError("pvdsp_render failed")
    return bufs
