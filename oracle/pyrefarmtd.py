"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libarmour_ref_armtd.so: the reference's ARMTD comparison planner (KPA sources compiled by
oracle/Makefile.ref, target `armtd`) behind oracle/ref_driver_armtd.cu.  Needs a GPU (the reference's collision kernels)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libarmour_ref_armtd.so")
_dp = C.POINTER(C.c_double)
NF = 7


def available():
    return os.path.exists(LIB)


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ReferenceArmtd:
    def __init__(self, q0, qd0, q_des, jrs, k_range, obstacles, nthreads=0):
        L = self.L = C.CDLL(LIB)
        L.refarmtd_build.restype = C.c_void_p
        L.refarmtd_build.argtypes = [_dp] * 6 + [C.c_int, C.c_int]
        for n in ("destroy", "num_constraints"):
            getattr(L, "refarmtd_" + n).argtypes = [C.c_void_p]
        L.refarmtd_bounds.argtypes = [C.c_void_p] + [_dp] * 4
        L.refarmtd_cost.argtypes = [C.c_void_p] + [_dp] * 3
        L.refarmtd_eval_g.argtypes = [C.c_void_p, _dp, _dp]
        L.refarmtd_eval_jac_g.argtypes = [C.c_void_p, _dp, _dp]
        L.refarmtd_finalize.argtypes = [C.c_void_p, _dp, _dp, C.c_double]
        L.refarmtd_link_sliced_center.argtypes = [C.c_void_p, _dp]
        L.refarmtd_link_gens.argtypes = [C.c_void_p, _dp]
        L.refarmtd_link_tables.argtypes = [C.c_void_p, _dp, C.c_int]
        self.T, self.NJ = L.refarmtd_num_time_steps(), L.refarmtd_num_joints()
        q0, qd0, q_des, jrs, k_range, obstacles = map(_f, (q0, qd0, q_des, jrs, k_range, obstacles))
        assert jrs.shape == (6, NF, self.T)
        self.O = obstacles.reshape(-1, 12).shape[0]
        self._h = L.refarmtd_build(_p(q0), _p(qd0), _p(q_des), _p(jrs), _p(k_range), _p(obstacles), self.O, nthreads)
        if not self._h:
            raise RuntimeError("the reference threw (or CUDA failed)")
        self.m = L.refarmtd_num_constraints(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.refarmtd_destroy(self._h)
            self._h = None

    def bounds(self):
        xl, xu, gl, gu = np.empty(NF), np.empty(NF), np.empty(self.m), np.empty(self.m)
        self.L.refarmtd_bounds(self._h, _p(xl), _p(xu), _p(gl), _p(gu))
        return gl, gu

    def cost(self, k):
        k = _f(k)
        f, grad = np.empty(1), np.empty(NF)
        self.L.refarmtd_cost(self._h, _p(k), _p(f), _p(grad))
        return float(f[0]), grad

    def eval_g(self, k):
        k, g = _f(k), np.empty(self.m)
        assert self.L.refarmtd_eval_g(self._h, _p(k), _p(g)) == 0
        return g

    def eval_jac_g(self, k):
        k, J = _f(k), np.zeros((self.m, NF))
        assert self.L.refarmtd_eval_jac_g(self._h, _p(k), _p(J)) == 0
        return J

    def finalize(self, k, g):
        k, g = _f(k), _f(g)
        return bool(self.L.refarmtd_finalize(self._h, _p(k), _p(g), 0.0))

    def link_sliced_center(self):
        out = np.empty((self.T, self.NJ, 3))
        self.L.refarmtd_link_sliced_center(self._h, _p(out))
        return out

    def link_gens(self):
        out = np.empty((self.T, self.NJ, 18))
        self.L.refarmtd_link_gens(self._h, _p(out))
        return out

    def link_tables(self):
        """list over (t, l) of (center[3], keys[n], coeff[n, 3])"""
        buf = np.empty(self.T * self.NJ * 4 * 200)
        n = self.L.refarmtd_link_tables(self._h, _p(buf), buf.size)
        assert n > 0
        out, k = [], 0
        for _ in range(self.T * self.NJ):
            cnt = int(buf[k])
            c = buf[k + 1:k + 4].copy()
            body = buf[k + 4:k + 4 + cnt * 4].reshape(cnt, 4)
            out.append((c, body[:, 0].astype(np.uint64), body[:, 1:].copy()))
            k += 4 + cnt * 4
        return out
