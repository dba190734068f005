"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes bindings of the two CPU sides of the robust-controller path:
  * OracleController — oracle/controller.cpp inside liboracle.so, the restatement;
  * ReferenceController — oracle/_ref/libarmour_ref_controller.so, the reference's own MEX/*.cpp sources compiled against the
    stand-in Eigen / Boost.Interval headers (oracle/Makefile.ref, target `mex`), the pin of the restatement.
Same method names and array conventions on both: one state per call, arrays of numJoints doubles.
Only tests/ and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MODEL = os.path.join(os.path.dirname(_HERE), "tests", "golden", "robot_models", "kinova_without_gripper.txt")
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Base:
    prefix = ""

    def __init__(self, lib, model_file=MODEL, eps=0.03):
        self.L = lib
        g = lambda name: getattr(lib, self.prefix + name)  # noqa: E731
        g("create").restype = C.c_void_p
        g("create").argtypes = [C.c_char_p, C.c_double]
        g("destroy").argtypes = [C.c_void_p]
        g("num_joints").argtypes = [C.c_void_p]
        g("rnea").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp]
        g("rnea_int").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
        g("int_model").argtypes = [C.c_void_p, _dp]
        self._g = g
        self._h = g("create")(model_file.encode(), eps)
        if not self._h:
            raise RuntimeError("cannot load robot model " + model_file)
        self.nj = g("num_joints")(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._g("destroy")(self._h)
            self._h = None

    def rnea(self, q, qd, qda, qdd, friction=False, gravity=True):
        q, qd, qda, qdd = map(_f, (q, qd, qda, qdd))
        tau = np.empty(self.nj)
        self._g("rnea")(self._h, _p(q), _p(qd), _p(qda), _p(qdd), int(friction), int(gravity), _p(tau))
        return tau

    def rnea_interval(self, q, qd, qda, qdd, friction=False, gravity=True):
        q, qd, qda, qdd = map(_f, (q, qd, qda, qdd))
        lo, hi = np.empty(self.nj), np.empty(self.nj)
        self._g("rnea_int")(self._h, _p(q), _p(qd), _p(qda), _p(qdd), int(friction), int(gravity), _p(lo), _p(hi))
        return lo, hi

    def interval_model(self):
        out = np.empty(self.nj * 74)
        self._g("int_model")(self._h, _p(out))
        return out.reshape(self.nj, 37, 2)


class OracleController(_Base):
    prefix = "orcctl_"

    def __init__(self, model_file=MODEL, eps=0.03):
        from oracle.pyoracle import build_lib
        lib = C.CDLL(build_lib())
        super().__init__(lib, model_file, eps)
        lib.orcctl_update.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, C.c_double, C.c_int] + [_dp] * 8

    def update(self, Kr, alpha, V_max, r_norm_threshold, q, qd, q_des, qd_des, qdd_des, friction=False):
        Kr, q, qd, q_des, qd_des, qdd_des = map(_f, (Kr, q, qd, q_des, qd_des, qdd_des))
        u, un, v = np.empty(self.nj), np.empty(self.nj), np.empty(self.nj)
        st = self.L.orcctl_update(self._h, _p(Kr), alpha, V_max, r_norm_threshold, int(friction), _p(q), _p(qd), _p(q_des),
                                  _p(qd_des), _p(qdd_des), _p(u), _p(un), _p(v))
        return u, un, v, st


REF_LIB = os.path.join(_HERE, "_ref", "libarmour_ref_controller.so")


def reference_available():
    return os.path.exists(REF_LIB)


class ReferenceController(_Base):
    prefix = "refctl_"

    def __init__(self, model_file=MODEL, eps=0.03):
        lib = C.CDLL(REF_LIB)
        super().__init__(lib, model_file, eps)
        lib.refctl_update.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double, C.c_double] + [_dp] * 8

    def update(self, Kr, alpha, V_max, r_norm_threshold, q, qd, q_des, qd_des, qdd_des):
        """as MEX/kinova_controller.cpp sets the controller up (no friction)"""
        Kr, q, qd, q_des, qd_des, qdd_des = map(_f, (Kr, q, qd, q_des, qd_des, qdd_des))
        u, un, v = np.empty(self.nj), np.empty(self.nj), np.empty(self.nj)
        st = self.L.refctl_update(self._h, _p(Kr), alpha, V_max, r_norm_threshold, _p(q), _p(qd), _p(q_des), _p(qd_des),
                                  _p(qdd_des), _p(u), _p(un), _p(v))
        return u, un, v, st
