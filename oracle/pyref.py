"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libarmour_ref.so: the REFERENCE's own PZsparse / BezierCurve /
KinematicsDynamics sources compiled against stand-in Eigen / Boost.Interval headers (oracle/Makefile.ref,
oracle/ref_driver.cpp).  Used to pin the restated oracle and to generate tests/golden/*.npz
(tools/make_golden.py).  Fixed to the reference's compile-time configuration (7 joints, 128 steps,
threshold 5e-4, k_range pi/48).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libarmour_ref.so")
REF_SRC = "/root/reference/kinova_src/kinova_simulator_interfaces/kinova_planner_realtime"
NF = 7
_LIB = None


def available(build: bool = True) -> bool:
    """True if the library exists (building it first when the reference sources are present)."""
    if build and os.path.isdir(REF_SRC):
        res = subprocess.run(["make", "-C", _HERE, "-f", "Makefile.ref"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + res.stdout + res.stderr)
    return os.path.exists(LIB_PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH)
        dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)
        L.ref_build.restype = C.c_void_p
        L.ref_build.argtypes = [dp, dp, dp, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_export.argtypes = [C.c_void_p, C.c_int, ip, dp, up, dp, C.c_int, ip, dp, up, dp, dp, dp, dp]
        L.ref_slice.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp, dp]
        L.ref_jrs.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class ReferenceProblem:
    def __init__(self, q0, qd0, qdd0, nthreads=0):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        q0, qd0, qdd0 = f(q0), f(qd0), f(qdd0)
        self._h = lib().ref_build(_dp(q0), _dp(qd0), _dp(qdd0), nthreads)
        if not self._h:
            raise RuntimeError("reference build threw")
        self.T, self.NJ = lib().ref_num_time_steps(), lib().ref_num_joints()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_destroy(self._h)
            self._h = None

    def tables(self, cap_link=64, cap_u=128):
        T, NJ = self.T, self.NJ
        r = dict(nl=np.zeros(T * NJ, np.int32), cl=np.zeros((T * NJ, 3)), hl=np.zeros((T * NJ, cap_link), np.uint64),
                 gl=np.zeros((T * NJ, cap_link, 3)), nu=np.zeros(T * NF, np.int32), cu=np.zeros(T * NF),
                 hu=np.zeros((T * NF, cap_u), np.uint64), gu=np.zeros((T * NF, cap_u)), ru=np.zeros(T * NF),
                 torque_radius=np.zeros((NF, T)), link_gens=np.zeros((T, NJ, 18)))
        ip, up = C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)
        mx = lib().ref_export(self._h, cap_link, r["nl"].ctypes.data_as(ip), _dp(r["cl"]), r["hl"].ctypes.data_as(up),
                              _dp(r["gl"]), cap_u, r["nu"].ctypes.data_as(ip), _dp(r["cu"]), r["hu"].ctypes.data_as(up),
                              _dp(r["gu"]), _dp(r["ru"]), _dp(r["torque_radius"]), _dp(r["link_gens"]))
        if mx < 0:
            raise RuntimeError(f"capacity exceeded ({-mx})")
        return r

    def slice(self, k):
        """Rows the reference computes on the host at k: torque rows + Jacobian, sliced link centres + d/dk,
        Bezier extremum rows + Jacobian."""
        k = np.ascontiguousarray(k, dtype=np.float64)
        T, NJ = self.T, self.NJ
        out = dict(g_torque=np.zeros(T * NF), jac_torque=np.zeros((T * NF, NF)), link_c=np.zeros((T, NJ, 3)),
                   dlink_c=np.zeros((T, NJ, NF, 3)), bez=np.zeros(4 * NF), dbez=np.zeros((4 * NF, NF)))
        lib().ref_slice(self._h, _dp(k), _dp(out["g_torque"]), _dp(out["jac_torque"]), _dp(out["link_c"]),
                        _dp(out["dlink_c"]), _dp(out["bez"]), _dp(out["dbez"]))
        return out
