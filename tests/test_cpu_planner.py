"""Host-side local solver (armour_b200/host/local_solver.cpp, the stand-in for Ipopt of the C++ host side, and the
algorithm the device solver k4 restates) checked on the CPU against an independent optimiser: it drives the ORACLE through
the TNLP callbacks (oracle/cpu_planner.cpp) and must end where scipy's SLSQP ends on the same f, g and Jacobian."""
import glob
import os

import numpy as np

from conftest import WORLDS


def test_local_solver_matches_slsqp_on_saved_worlds(built):
    from scipy.optimize import minimize

    from armour_b200 import worlds
    from oracle.pyoracle import OracleProblem
    for path in sorted(glob.glob(os.path.join(WORLDS, "scene_*.csv")))[:4]:
        q0, qd0, qdd0, q_des, obs = worlds.config1_problem(path)
        orc = OracleProblem().build(q0, qd0, qdd0, obs)
        gl, gu = orc.bounds()
        fin = gl > -1e18

        def cons(x):
            g = orc.eval_g(x)
            return np.concatenate([g[fin] - gl[fin], gu - g])

        def cjac(x):
            J = orc.eval_jac_g(x)
            return np.vstack([J[fin], -J])

        r = minimize(lambda x: orc.cost(q_des, x), np.zeros(7), jac=lambda x: orc.cost_grad(q_des, x),
                     bounds=[(-1, 1)] * 7, constraints=[{"type": "ineq", "fun": cons, "jac": cjac}], method="SLSQP",
                     options={"maxiter": 200, "ftol": 1e-12})
        ok_ref, _ = orc.verdict(orc.eval_g(r.x))
        k, ok, first, iters = orc.solve(q_des)
        assert ok == ok_ref
        if ok_ref:
            assert orc.cost(q_des, k) <= r.fun + 1e-6
            assert orc.verdict(orc.eval_g(k))[0]
            assert 1 <= iters <= 60
