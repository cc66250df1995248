"""planeverb_b200 -- B200-native (CUDA sm_100a) replacement for Planeverb's hot path: the 2-D FDTD
acoustic solve and the per-cell impulse-response analyzer, behind Planeverb's own C++/C API.

Layout: csrc/ (CUDA kernels + C-ABI + host C++), lib/ (built shared library, git-ignored),
pvcuda.py (ctypes mirror of include/planeverb_cuda.h + planeverb_ext.h), scenes/ (the reference's
.pv scene fixtures as JSON)."""
__all__ = ["pvcuda"]
