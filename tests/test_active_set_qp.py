"""The exact QP solver of the SQP step (armour_b200/host/active_set_qp.h — shared by the host solver and the device solver
k4) against independent checks: the KKT conditions of its answer (non-negative multipliers from scipy's NNLS) and scipy's
SLSQP on the same QP."""
import ctypes as C

import numpy as np


def _qp(n, h, c, A, b):
    from oracle.pyoracle import lib
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_qp.argtypes = [C.c_int, C.c_double, dp, C.c_int, dp, dp, dp]
    c, A, b = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, A, b))
    d = np.empty(n)
    it = L.orc_qp(n, h, c.ctypes.data_as(dp), A.shape[0], A.ctypes.data_as(dp), b.ctypes.data_as(dp), d.ctypes.data_as(dp))
    return d, it


def _kkt_residual(h, c, A, b, d):
    """distance of -(h d + c) from the cone of the active normals (0 at the optimum)"""
    from scipy.optimize import nnls
    act = A @ d - b > -1e-9
    if not act.any():
        return float(np.linalg.norm(h * d + c))
    _, res = nnls(A[act].T, -(h * d + c))
    return float(res)


def test_random_feasible_qps_satisfy_kkt_and_match_slsqp(built):
    from scipy.optimize import minimize
    rng = np.random.default_rng(7)
    for trial in range(200):
        n = 7
        m = int(rng.integers(1, 120))
        h = float(rng.uniform(0.1, 50))
        c = rng.normal(size=n) * rng.uniform(0.1, 20)
        A = rng.normal(size=(m, n)) * rng.uniform(0.01, 100, size=(m, 1))
        if trial % 3 == 0:  # clusters of nearly parallel rows, as neighbouring time intervals give
            A[1::2] = A[:-1:2][: A[1::2].shape[0]] * (1 + 1e-7 * rng.normal(size=(A[1::2].shape[0], n)))
        x_in = rng.normal(size=n) * 0.2
        b = A @ x_in + rng.uniform(0.0, 1.0, size=m) * np.linalg.norm(A, axis=1)  # x_in is strictly inside
        box = np.vstack([np.eye(n), -np.eye(n)])
        A = np.vstack([A, box])
        b = np.concatenate([b, np.full(2 * n, 1.0)])
        d, it = _qp(n, h, c, A, b)
        scale = np.linalg.norm(A, axis=1)
        assert np.max((A @ d - b) / scale) <= 1e-9, f"trial {trial}: infeasible answer"
        assert _kkt_residual(h, c, A, b, d) <= 1e-7 * (1 + np.linalg.norm(c)), f"trial {trial}: not a KKT point"
        assert 0 <= it < 200
        if trial % 10 == 0:
            r = minimize(lambda x: 0.5 * h * x @ x + c @ x, x_in, jac=lambda x: h * x + c, method="SLSQP",
                         constraints=[{"type": "ineq", "fun": lambda x: b - A @ x, "jac": lambda x: -A}],
                         options={"maxiter": 500, "ftol": 1e-14})
            if r.success:
                assert 0.5 * h * d @ d + c @ d <= r.fun + 1e-7 * (1 + abs(r.fun))


def test_unconstrained_and_single_row(built):
    d, it = _qp(7, 2.0, np.arange(7.0), np.zeros((1, 7)), np.ones(1))
    assert np.allclose(d, -np.arange(7.0) / 2.0) and it == 0
    a = np.zeros((1, 7))
    a[0, 0] = 1.0
    d, it = _qp(7, 1.0, -np.ones(7), a, np.array([0.25]))  # minimiser (1,..,1) cut by x0 <= 0.25
    assert np.allclose(d, [0.25] + [1.0] * 6) and it == 1


def test_contradicting_rows_end_with_the_consistent_subset(built):
    """an infeasible linearisation (x0 <= -1 and x0 >= 1): the solver must stop, keeping the first (most violated) row"""
    A = np.zeros((2, 7))
    A[0, 0], A[1, 0] = 1.0, -1.0
    d, it = _qp(7, 1.0, np.zeros(7), A, np.array([-1.0, -2.0]))
    assert it < 200 and np.all(np.isfinite(d))
    assert abs(d[0] - 2.0) < 1e-12  # row 1 (violation 2) is taken first: x0 >= 2; row 0 then contradicts it
