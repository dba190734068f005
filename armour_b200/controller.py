"""Host-side mirror of the reference's MEX entry `kinova_controller` (MEX/kinova_controller.cpp), batched over states.

    [u, tau, v] = kinova_controller(Kr, alpha, V_max, r_norm_threshold, q, qd, q_des, qd_des, qdd_des [, eps])

becomes `RobustController(model_file, eps).update(Kr, alpha, V_max, r_norm_threshold, q, qd, q_des, qd_des, qdd_des)` with
[n, numJoints] arrays; `rnea` / `rnea_interval` are passRNEA / passRNEA_Int (MEX/rnea.cpp).  All of it runs in
libarmour_b200.so (csrc/controller.cuh); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArmourError, ControllerGains, dp, ip


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(dp)


class RobustController:
    def __init__(self, model_file: str, model_uncertainty: float = 0.03, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.armour_controller_create(model_file.encode(), float(model_uncertainty), int(device), C.byref(h))
        if rc != 0:
            raise ArmourError(rc, self.lib.armour_status_string(rc).decode() + ": " +
                              self.lib.armour_controller_create_error().decode())
        self._h = h
        self.nj = self.lib.armour_controller_num_joints(h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.armour_controller_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise ArmourError(rc, self.lib.armour_status_string(rc).decode() + ": " +
                              self.lib.armour_controller_last_error(self._h).decode())

    @property
    def kernel_launches(self):
        return int(self.lib.armour_controller_kernel_launches(self._h))

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.armour_controller_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.armour_controller_synchronize(self._h))

    def interval_model(self):
        out = np.empty(self.nj * 74)
        self._check(self.lib.armour_controller_get_interval_model(self._h, _dp(out)))
        return out.reshape(self.nj, 37, 2)

    def _states(self, *arrays):
        arrs = [np.atleast_2d(_f64(a)) for a in arrays]
        n = arrs[0].shape[0]
        for a in arrs:
            if a.shape != (n, self.nj):
                raise ValueError(f"state arrays must be [n, {self.nj}]")
        return n, arrs

    def rnea(self, q, qd, qda, qdd, friction=False, gravity=True, nominal=True, interval=True):
        """passRNEA and / or passRNEA_Int of n states: returns (tau, tau_lo, tau_hi), None where not requested."""
        n, (q, qd, qda, qdd) = self._states(q, qd, qda, qdd)
        tau = np.empty((n, self.nj)) if nominal else None
        lo = np.empty((n, self.nj)) if interval else None
        hi = np.empty((n, self.nj)) if interval else None
        self._check(self.lib.armour_controller_rnea(self._h, n, _dp(q), _dp(qd), _dp(qda), _dp(qdd), int(friction), int(gravity),
                                                    _dp(tau) if nominal else None, _dp(lo) if interval else None,
                                                    _dp(hi) if interval else None))
        return tau, lo, hi

    def rnea_device(self, n, d_q, d_qd, d_qda, d_qdd, d_tau=0, d_tau_lo=0, d_tau_hi=0, d_sincos=0, friction=False, gravity=True):
        """device pointers (ints); asynchronous on the controller's stream"""
        v = lambda p: C.c_void_p(p) if p else None  # noqa: E731
        self._check(self.lib.armour_controller_rnea_device(self._h, int(n), v(d_q), v(d_qd), v(d_qda), v(d_qdd), v(d_sincos),
                                                           int(friction), int(gravity), v(d_tau), v(d_tau_lo), v(d_tau_hi)))

    def _gains(self, Kr, alpha, V_max, r_norm_threshold, friction):
        Kr = _f64(Kr)
        if Kr.ndim == 2:  # the MEX entry takes the diagonal; accept the matrix too
            Kr = _f64(np.diag(Kr))
        if Kr.shape != (self.nj,):
            raise ValueError(f"Kr must have {self.nj} entries")
        g = ControllerGains(_dp(Kr), float(alpha), float(V_max), float(r_norm_threshold), int(friction))
        return g, Kr  # keep Kr alive

    def update(self, Kr, alpha, V_max, r_norm_threshold, q, qd, q_des, qd_des, qdd_des, friction=False):
        """RobustController::update (ARMOUR method) of n states: returns (u, u_nominal, v, status)."""
        n, (q, qd, q_des, qd_des, qdd_des) = self._states(q, qd, q_des, qd_des, qdd_des)
        g, keep = self._gains(Kr, alpha, V_max, r_norm_threshold, friction)
        u, un, v = (np.empty((n, self.nj)) for _ in range(3))
        st = np.empty(n, dtype=np.int32)
        self._check(self.lib.armour_controller_update(self._h, n, C.byref(g), _dp(q), _dp(qd), _dp(q_des), _dp(qd_des), _dp(qdd_des),
                                                      _dp(u), _dp(un), _dp(v), st.ctypes.data_as(ip)))
        return u, un, v, st

    def update_device(self, n, Kr, alpha, V_max, r_norm_threshold, d_q, d_qd, d_q_des, d_qd_des, d_qdd_des, d_u, d_u_nominal=0,
                      d_v=0, d_status=0, d_sincos=0, friction=False):
        g, keep = self._gains(Kr, alpha, V_max, r_norm_threshold, friction)
        v = lambda p: C.c_void_p(p) if p else None  # noqa: E731
        self._check(self.lib.armour_controller_update_device(self._h, int(n), C.byref(g), v(d_q), v(d_qd), v(d_q_des), v(d_qd_des),
                                                             v(d_qdd_des), v(d_sincos), v(d_u), v(d_u_nominal), v(d_v), v(d_status)))
