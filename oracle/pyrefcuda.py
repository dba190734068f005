"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libarmour_ref_cuda.so: the REFERENCE's complete planner path — PZsparse.cu,
Trajectory.cu, Dynamics.cu, CollisionChecking.cu (its CUDA kernels) and NLPclass.cu (armtd_NLP) — compiled by nvcc
from /root/reference with the reference's flags (oracle/Makefile.ref `cuda`, oracle/ref_driver_nlp.cu).  It needs a
GPU, so it only runs on the GPU box: tools/make_golden_collision.py freezes its outputs as tests/golden/refcuda/*.npz
and tests/test_refcuda_gpu.py compares live when the library travelled.  Fixed to the reference's compile-time
configuration (7 joints, 128 steps, threshold 5e-4, k_range pi/48, at most 40 obstacles).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libarmour_ref_cuda.so")
NF = 7
_LIB = None


def available() -> bool:
    """True if the library was built (here, from /root/reference) and a CUDA device answers."""
    if not os.path.exists(LIB_PATH):
        return False
    try:
        return lib().reffull_cuda_devices() > 0
    except OSError:
        return False


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.reffull_build.restype = C.c_void_p
        L.reffull_build.argtypes = [dp, dp, dp, dp, dp, C.c_int, C.c_int]
        L.reffull_destroy.argtypes = [C.c_void_p]
        L.reffull_num_constraints.argtypes = [C.c_void_p]
        L.reffull_bounds.argtypes = [C.c_void_p, dp, dp, dp, dp]
        L.reffull_cost.argtypes = [C.c_void_p, dp, dp, dp]
        L.reffull_eval_g.argtypes = [C.c_void_p, dp, dp]
        L.reffull_eval_jac_g.argtypes = [C.c_void_p, dp, dp]
        L.reffull_finalize.argtypes = [C.c_void_p, dp, dp, C.c_double]
        L.reffull_link_sliced_center.argtypes = [C.c_void_p, dp]
        L.reffull_hyperplanes.argtypes = [C.c_void_p, dp, dp, dp]
        L.reffull_problem.restype = C.c_void_p
        L.reffull_problem.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ReferencePlanner:
    """armtd_NLP of the reference on one planning problem (Obstacles + reach sets + NLP object)."""

    T, NJ = 128, 7

    def __init__(self, q0, qd0, qdd0, q_des, obstacles, nthreads=0):
        obs = _f64(obstacles).reshape(-1, 12)
        self.nobs = obs.shape[0]
        a, b, c, d = _f64(q0), _f64(qd0), _f64(qdd0), _f64(q_des)
        self._h = lib().reffull_build(_dp(a), _dp(b), _dp(c), _dp(d), _dp(obs), self.nobs, nthreads)
        if not self._h:
            raise RuntimeError("the reference threw (too many obstacles?) or CUDA failed")
        self.m = lib().reffull_num_constraints(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().reffull_destroy(self._h)
            self._h = None

    def bounds(self):
        xl, xu, gl, gu = np.empty(NF), np.empty(NF), np.empty(self.m), np.empty(self.m)
        lib().reffull_bounds(self._h, _dp(xl), _dp(xu), _dp(gl), _dp(gu))
        return xl, xu, gl, gu

    def cost(self, k):
        k = _f64(k)
        f, grad = np.empty(1), np.empty(NF)
        lib().reffull_cost(self._h, _dp(k), _dp(f), _dp(grad))
        return float(f[0]), grad

    def eval_g(self, k):
        k = _f64(k)
        g = np.empty(self.m)
        if lib().reffull_eval_g(self._h, _dp(k), _dp(g)) != 0:
            raise RuntimeError("CUDA error in the reference's eval_g")
        return g

    def eval_jac_g(self, k):
        k = _f64(k)
        v = np.empty((self.m, NF))
        if lib().reffull_eval_jac_g(self._h, _dp(k), _dp(v)) != 0:
            raise RuntimeError("CUDA error in the reference's eval_jac_g")
        return v

    def finalize(self, k, g, obj=0.0):
        k, g = _f64(k), _f64(g)
        return bool(lib().reffull_finalize(self._h, _dp(k), _dp(g), obj))

    def link_sliced_center(self):
        out = np.empty((self.T, self.NJ, 3))
        lib().reffull_link_sliced_center(self._h, _dp(out))
        return out

    def hyperplanes(self):
        n = self.T * self.NJ * self.nobs * 36
        A, d, delta = np.empty((n, 3)), np.empty(n), np.empty(n)
        if lib().reffull_hyperplanes(self._h, _dp(A), _dp(d), _dp(delta)) != 0:
            raise RuntimeError("CUDA error copying the reference's half-spaces")
        shp = (self.T, self.NJ, self.nobs, 36)
        return A.reshape(shp + (3,)), d.reshape(shp), delta.reshape(shp)
