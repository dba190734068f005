"""ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources: reach-set build, slices and Bezier rows through oracle/_ref/libarmour_ref.so (tests/test_oracle_pinned.py); collision rows, bounds, cost and verdict through the reference's CUDA kernels and armtd_NLP compiled by nvcc (oracle/_ref/libarmour_ref_cuda.so, run on a B200, frozen as tests/golden/refcuda/, tests/test_refcuda_golden.py).

ctypes binding of oracle/liboracle.so (the CPU restatement of the reference planner's hot path).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under armour_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NF = 7
JRS_FIELDS = ("cos_center", "cos_k", "cos_e", "sin_center", "sin_k", "sin_e", "qd_center", "qd_k", "qd_e", "qda_e",
              "qdd_center", "qdd_k", "qdd_e")


def build_lib(force: bool = False) -> str:
    path = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(path):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return path


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_lib())
        dp = C.POINTER(C.c_double)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.c_int, C.c_double, C.c_double]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_build.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_int, C.c_int]
        for name in ("orc_num_constraints", "orc_num_joints", "orc_num_time_steps"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_build_ms.restype = C.c_double
        L.orc_build_ms.argtypes = [C.c_void_p]
        L.orc_eval_g.argtypes = [C.c_void_p, dp, dp]
        L.orc_eval_jac_g.argtypes = [C.c_void_p, dp, dp]
        L.orc_bounds.argtypes = [C.c_void_p, dp, dp]
        L.orc_verdict.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int)]
        L.orc_cost.restype = C.c_double
        L.orc_cost.argtypes = [C.c_void_p, dp, dp]
        L.orc_cost_grad.argtypes = [C.c_void_p, dp, dp, dp]
        for name in ("orc_get_torque_radius", "orc_get_link_gens", "orc_get_link_sliced_center", "orc_get_jrs"):
            getattr(L, name).argtypes = [C.c_void_p, dp]
        L.orc_get_hyperplanes.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_get_stats.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.orc_solve.argtypes = [C.c_void_p, dp, C.c_int, C.c_double, dp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]
        ip = C.POINTER(C.c_int)
        up = C.POINTER(C.c_ulonglong)
        L.orc_export_reachsets.argtypes = [C.c_void_p, C.c_int, ip, dp, up, dp, C.c_int, ip, dp, up, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleProblem:
    """One planning problem evaluated by the CPU oracle (mirrors armour_main.cu + armtd_NLP)."""

    def __init__(self, model_id=0, num_time_steps=128, simplify_threshold=5e-4, k_range=None, max_obstacles=40,
                 mass_uncertainty=-1.0, inertia_uncertainty=-1.0):
        kr = None if k_range is None else _f64(k_range)
        self._h = lib().orc_create(model_id, num_time_steps, simplify_threshold, None if kr is None else _dp(kr),
                                   max_obstacles, mass_uncertainty, inertia_uncertainty)
        self.nobs = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def build(self, q0, qd0, qdd0, obstacles, nthreads=0):
        q0, qd0, qdd0 = _f64(q0), _f64(qd0), _f64(qdd0)
        obs = _f64(obstacles).reshape(-1, 12)
        self.nobs = obs.shape[0]
        rc = lib().orc_build(self._h, _dp(q0), _dp(qd0), _dp(qdd0), _dp(obs), self.nobs, nthreads)
        if rc != 0:
            raise RuntimeError("oracle build failed (too many obstacles?)")
        self.T = lib().orc_num_time_steps(self._h)
        self.NJ = lib().orc_num_joints(self._h)
        self.m = lib().orc_num_constraints(self._h)
        return self

    @property
    def build_ms(self):
        return lib().orc_build_ms(self._h)

    def eval_g(self, k):
        k = _f64(k)
        g = np.empty(self.m)
        lib().orc_eval_g(self._h, _dp(k), _dp(g))
        return g

    def eval_jac_g(self, k):
        k = _f64(k)
        v = np.empty((self.m, NF))
        lib().orc_eval_jac_g(self._h, _dp(k), _dp(v))
        return v

    def bounds(self):
        gl, gu = np.empty(self.m), np.empty(self.m)
        lib().orc_bounds(self._h, _dp(gl), _dp(gu))
        return gl, gu

    def verdict(self, g):
        g = _f64(g)
        first = C.c_int(-1)
        ok = lib().orc_verdict(self._h, _dp(g), C.byref(first))
        return bool(ok), first.value

    def cost(self, q_des, k):
        q_des, k = _f64(q_des), _f64(k)
        return lib().orc_cost(self._h, _dp(q_des), _dp(k))

    def cost_grad(self, q_des, k):
        q_des, k = _f64(q_des), _f64(k)
        out = np.empty(NF)
        lib().orc_cost_grad(self._h, _dp(q_des), _dp(k), _dp(out))
        return out

    def solve(self, q_des, max_iter=0, max_wall_time=0.0):
        """Plan on the CPU: the product's host-side local solver (stand-in for Ipopt) driving this oracle through the
        TNLP callbacks (oracle/cpu_planner.cpp).  Returns (k_opt, feasible, first violated row, iterations)."""
        q_des = _f64(q_des)
        k = np.empty(NF)
        ok, first, it = C.c_int(0), C.c_int(-1), C.c_int(0)
        lib().orc_solve(self._h, _dp(q_des), max_iter, max_wall_time, _dp(k), C.byref(ok), C.byref(first), C.byref(it))
        return k, bool(ok.value), first.value, it.value

    def torque_radius(self):
        out = np.empty((NF, self.T))
        lib().orc_get_torque_radius(self._h, _dp(out))
        return out

    def link_gens(self):
        """[T, NJ, 3, 6] (row, column) view of the column-major 3x6 generator matrices."""
        out = np.empty((self.T, self.NJ, 6, 3))
        lib().orc_get_link_gens(self._h, _dp(out))
        return out.transpose(0, 1, 3, 2)

    def link_sliced_center(self):
        out = np.empty((self.T, self.NJ, 3))
        lib().orc_get_link_sliced_center(self._h, _dp(out))
        return out

    def hyperplanes(self):
        n = self.T * self.NJ * self.nobs * 36
        A, d, delta = np.empty((n, 3)), np.empty(n), np.empty(n)
        lib().orc_get_hyperplanes(self._h, _dp(A), _dp(d), _dp(delta))
        shp = (self.T, self.NJ, self.nobs, 36)
        return A.reshape(shp + (3,)), d.reshape(shp), delta.reshape(shp)

    def jrs(self):
        out = np.empty((NF, self.T, len(JRS_FIELDS)))
        lib().orc_get_jrs(self._h, _dp(out))
        return {f: out[:, :, i] for i, f in enumerate(JRS_FIELDS)}

    def stats(self):
        out = (C.c_ulonglong * 8)()
        lib().orc_get_stats(self._h, out)
        keys = ("n_simplify", "n_mul", "n_pairs", "flops", "terms_sorted", "max_terms", "max_monos", "near_threshold")
        return dict(zip(keys, [int(x) for x in out]))

    def export_reachsets(self, cap_link=64, cap_u=128):
        T, NJ = self.T, self.NJ
        r = dict(
            nl=np.zeros(T * NJ, np.int32), cl=np.zeros((T * NJ, 3)), hl=np.zeros((T * NJ, cap_link), np.uint64),
            gl=np.zeros((T * NJ, cap_link, 3)), nu=np.zeros(T * NF, np.int32), cu=np.zeros(T * NF),
            hu=np.zeros((T * NF, cap_u), np.uint64), gu=np.zeros((T * NF, cap_u)), ru=np.zeros(T * NF))
        ip, up = C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)
        mx = lib().orc_export_reachsets(
            self._h, cap_link, r["nl"].ctypes.data_as(ip), _dp(r["cl"]), r["hl"].ctypes.data_as(up), _dp(r["gl"]),
            cap_u, r["nu"].ctypes.data_as(ip), _dp(r["cu"]), r["hu"].ctypes.data_as(up), _dp(r["gu"]), _dp(r["ru"]))
        if mx < 0:
            raise RuntimeError(f"reach-set export capacity exceeded (largest table {-mx})")
        r["max_monos"] = mx
        return r

    def tables(self, cap_link=64, cap_u=128):
        """k-only reach-set tables in the neutral layout of armour_import_reachsets (include/armour_b200.h)."""
        r = self.export_reachsets(cap_link, cap_u)
        r["torque_radius"] = np.ascontiguousarray(self.torque_radius())
        raw = np.empty((self.T, self.NJ, 18))
        lib().orc_get_link_gens(self._h, _dp(raw))
        r["link_gens"] = raw
        return r


class OracleArmtd(OracleProblem):
    """The ARMTD comparison planner (KPA) evaluated by the CPU oracle (oracle/armtd.cpp): constant-acceleration trajectory,
    offline joint reachable set handed in, forward kinematics only, 100 time steps."""

    def __init__(self, simplify_threshold=5e-4, max_obstacles=40, num_time_steps=100):
        super().__init__(0, num_time_steps, simplify_threshold, None, max_obstacles)
        L = lib()
        L.orc_armtd_build.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 5 + [C.c_int, C.c_int]
        L.orc_armtd_num_constraints.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        L.orc_armtd_eval_g.argtypes = [C.c_void_p, dp, dp]
        L.orc_armtd_eval_jac_g.argtypes = [C.c_void_p, dp, dp]
        L.orc_armtd_bounds.argtypes = [C.c_void_p, dp, dp]
        L.orc_armtd_verdict.argtypes = [C.c_void_p, dp, C.POINTER(C.c_int)]
        L.orc_armtd_cost.restype = C.c_double
        L.orc_armtd_cost.argtypes = [C.c_void_p, dp, dp, dp]

    def build(self, q0, qd0, jrs, k_range, obstacles, nthreads=0):
        q0, qd0, jrs, k_range = _f64(q0), _f64(qd0), _f64(jrs), _f64(k_range)
        obs = _f64(obstacles).reshape(-1, 12)
        self.nobs = obs.shape[0]
        self.T = lib().orc_num_time_steps(self._h)
        assert jrs.shape == (6, NF, self.T)
        if lib().orc_armtd_build(self._h, _dp(q0), _dp(qd0), _dp(jrs), _dp(k_range), _dp(obs), self.nobs, nthreads) != 0:
            raise RuntimeError("oracle build failed (too many obstacles?)")
        self.NJ = lib().orc_num_joints(self._h)
        self.m = lib().orc_armtd_num_constraints(self._h)
        return self

    def eval_g(self, k):
        k, g = _f64(k), np.empty(self.m)
        lib().orc_armtd_eval_g(self._h, _dp(k), _dp(g))
        return g

    def eval_jac_g(self, k):
        k, v = _f64(k), np.empty((self.m, NF))
        lib().orc_armtd_eval_jac_g(self._h, _dp(k), _dp(v))
        return v

    def bounds(self):
        gl, gu = np.empty(self.m), np.empty(self.m)
        lib().orc_armtd_bounds(self._h, _dp(gl), _dp(gu))
        return gl, gu

    def verdict(self, g):
        g = _f64(g)
        first = C.c_int(-1)
        ok = lib().orc_armtd_verdict(self._h, _dp(g), C.byref(first))
        return bool(ok), first.value

    def cost(self, q_des, k):
        q_des, k = _f64(q_des), _f64(k)
        return lib().orc_armtd_cost(self._h, _dp(q_des), _dp(k), None)

    def cost_grad(self, q_des, k):
        q_des, k, grad = _f64(q_des), _f64(k), np.empty(NF)
        lib().orc_armtd_cost(self._h, _dp(q_des), _dp(k), _dp(grad))
        return grad

    def link_tables(self, cap_link=64):
        r = self.export_reachsets(cap_link, 8)
        return r["nl"], r["cl"], r["hl"], r["gl"]

    def solve(self, q_des, max_iter=0, tol=1e-7):
        """Plan on the CPU (oracle/cpu_planner.cpp: the product's host-side local solver over this oracle, KPA's tolerance).
        Returns (k_opt, feasible, first violated row, iterations); link_sliced_center() then holds the centres at k_opt."""
        q_des, k = _f64(q_des), np.empty(NF)
        ok, first, it = C.c_int(0), C.c_int(-1), C.c_int(0)
        L = lib()
        L.orc_armtd_solve.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(C.c_double),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_armtd_solve(self._h, _dp(q_des), max_iter, tol, _dp(k), C.byref(ok), C.byref(first), C.byref(it))
        return k, bool(ok.value), first.value, it.value
