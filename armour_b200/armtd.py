"""Host-side mirror of the reference's ARMTD comparison planner (kinova_planner_realtime_armtd_comparison: armtd_main.cu +
armtd_NLP) over the `armour_armtd_*` section of the C ABI: same method names as the TNLP members the reference's Ipopt run
calls.  All compute is in libarmour_b200.so; there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NF, ArmourError, Config, dp, ip

T = 100  # KPA/Parameters.h:17


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(dp)


class ArmtdPlanner:
    def __init__(self, max_obstacles=40, simplify_threshold=5e-4, device=0):
        self.lib = _lib.load()
        cfg = Config()
        self.lib.armour_config_default(C.byref(cfg))
        cfg.max_obstacles = int(max_obstacles)
        cfg.simplify_threshold = float(simplify_threshold)
        cfg.device = int(device)
        h = C.c_void_p()
        rc = self.lib.armour_armtd_ctx_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ArmourError(rc, self.lib.armour_status_string(rc).decode())
        self._h = h
        self.NJ = self.lib.armour_num_joints(h)
        self.m = 0

    def close(self):
        if getattr(self, "_h", None):
            self.lib.armour_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise ArmourError(rc, self.lib.armour_status_string(rc).decode() + ": " + self.lib.armour_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return int(self.lib.armour_kernel_launches(self._h))

    def build(self, q0, qd0, jrs, k_range, obstacles):
        """jrs: [6, 7, 100] = c_cos, g_cos, r_cos, c_sin, g_sin, r_sin (the arrays of KPA's input file)"""
        q0, qd0, jrs, k_range = _f64(q0), _f64(qd0), _f64(jrs), _f64(k_range)
        if jrs.shape != (6, NF, T):
            raise ValueError("jrs must be [6, 7, 100]")
        obs = _f64(obstacles).reshape(-1, 12)
        self._check(self.lib.armour_armtd_build(self._h, _dp(q0), _dp(qd0), _dp(jrs), _dp(k_range), _dp(obs), obs.shape[0]))
        self.m = self.lib.armour_armtd_num_constraints(self._h)
        return self

    def eval_g(self, k):
        k, g = _f64(k), np.empty(self.m)
        self._check(self.lib.armour_armtd_eval(self._h, _dp(k), _dp(g), None))
        return g

    def eval_jac_g(self, k):
        k, v = _f64(k), np.empty((self.m, NF))
        self._check(self.lib.armour_armtd_eval(self._h, _dp(k), None, _dp(v)))
        return v

    def eval(self, k):
        k, g, v = _f64(k), np.empty(self.m), np.empty((self.m, NF))
        self._check(self.lib.armour_armtd_eval(self._h, _dp(k), _dp(g), _dp(v)))
        return g, v

    def get_bounds_info(self):
        gl, gu = np.empty(self.m), np.empty(self.m)
        self._check(self.lib.armour_armtd_get_bounds(self._h, _dp(gl), _dp(gu)))
        return gl, gu

    def finalize_solution(self, g):
        g = _f64(g)
        ok, first = C.c_int(0), C.c_int(-1)
        self._check(self.lib.armour_armtd_verdict(self._h, _dp(g), C.byref(ok), C.byref(first)))
        return bool(ok.value), first.value

    def eval_f(self, q_des, k):
        q_des, k = _f64(q_des), _f64(k)
        f = C.c_double(0)
        self._check(self.lib.armour_armtd_cost(self._h, _dp(q_des), _dp(k), C.cast(C.byref(f), dp), None))
        return f.value

    def eval_grad_f(self, q_des, k):
        q_des, k, grad = _f64(q_des), _f64(k), np.empty(NF)
        self._check(self.lib.armour_armtd_cost(self._h, _dp(q_des), _dp(k), None, _dp(grad)))
        return grad

    def link_sliced_center(self):
        out = np.empty((T, self.NJ, 3))
        self._check(self.lib.armour_armtd_get_link_sliced_center(self._h, _dp(out)))
        return out

    def link_independent_generators(self):
        out = np.empty((T, self.NJ, 18))
        self._check(self.lib.armour_armtd_get_link_independent_generators(self._h, _dp(out)))
        return out
