"""Parity on the larger BASELINE.json configurations: config 3 = 40 obstacles with the 8-link
gripper model and 10 % inertial uncertainty; config 4 = 100 obstacles (beyond the reference's MAX_OBSTACLE_NUM
of 40, KPR/Parameters.h:26) and a lower simplify threshold (more monomials everywhere)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K_TEST = np.array([0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7])


def _check(eng, ref, ks, tol=1e-9):
    assert eng.m == ref.m
    for k in ks:
        g, jac = eng.eval(k)
        g_ref, j_ref = ref.eval_g(k), ref.eval_jac_g(k)
        assert np.max(np.abs(g[0] - g_ref)) <= tol
        assert np.max(np.abs(jac[0] - j_ref)) <= tol
        assert eng.finalize_solution(g[0]) == ref.verdict(g_ref)
    tr, tr_ref = eng.torque_radius()[0], ref.torque_radius()
    assert np.all(tr >= tr_ref) and np.max((tr - tr_ref) / tr_ref) <= 1e-10


def test_config3_gripper_40_obstacles_uncertain_payload(built):
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, _, obs = worlds.random_problems(1, 40, seed=3)
    kw = dict(mass_uncertainty=0.10, inertia_uncertainty=0.10)
    ref = OracleProblem(model_id=1, max_obstacles=40, **kw).build(q0[0], qd0[0], qdd0[0], obs[0])
    eng = ReachSetEngine(max_problems=1, max_obstacles=40, robot_model=1, cap_link=64, cap_torque=128, **kw)
    eng.build(q0[0], qd0[0], qdd0[0], obs[0])
    assert eng.NJ == 8 and eng.m == 7 * 128 + 8 * 128 * 40 + 28
    _check(eng, ref, [np.zeros(7), K_TEST])


@pytest.mark.parametrize("thr", [5e-4, 2e-4, 5e-5, 5e-6])
def test_config4_100_obstacles_lower_threshold(built, thr):
    """BASELINE config 4: 100 obstacles, SIMPLIFY_THRESHOLD down to 5e-6, where intermediates reach 4 403
    monomials and a cross product 113 121 terms on this problem (oracle statistics) — beyond the 16-bit term lists, the
    survivor masks of the control block (8 192 candidates per merge) and the register path of the block sort (512
    survivors): the accumulator-table path of the cross product, the mask behind the scratch and the scatter sort through
    the global pool are exercised."""
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, _, obs = worlds.random_problems(1, 100, seed=4)
    ref = OracleProblem(simplify_threshold=thr, max_obstacles=100).build(q0[0], qd0[0], qdd0[0], obs[0])
    from armour_b200 import ArmourError
    eng = None
    for cap_work in ((16384, 32768) if thr < 5e-5 else (4096,)):
        # at 5e-6 the F / N blocks kept for the backward pass hold ~4 000 monomials each: the work capacity that fits them
        # is found here, and a capacity that does not must fail loudly (ARMOUR_ERR_CAPACITY), never truncate
        eng = ReachSetEngine(max_problems=1, max_obstacles=100, simplify_threshold=thr, cap_link=128, cap_torque=256,
                             cap_work=cap_work)
        try:
            eng.build(q0[0], qd0[0], qdd0[0], obs[0])
            break
        except ArmourError as exc:
            assert exc.code == -4
            eng.close()
            eng = None
    assert eng is not None, "no tested work capacity fits threshold %g" % thr
    assert eng.m == 7 * 128 + 7 * 128 * 100 + 28
    _check(eng, ref, [K_TEST, -K_TEST])


def test_too_small_capacity_is_reported_not_truncated(built):
    """A monomial table that does not fit its configured capacity fails the build with ARMOUR_ERR_CAPACITY, and the
    problem can never be taken for feasible afterwards: evaluations return fail-safe rows."""
    from armour_b200 import ArmourError, ReachSetEngine, worlds
    q0, qd0, qdd0, q_des, obs = worlds.random_problems(1, 2, seed=8)
    eng = ReachSetEngine(max_problems=1, max_obstacles=2, cap_link=8, cap_torque=8)
    with pytest.raises(ArmourError) as ei:
        eng.build(q0[0], qd0[0], qdd0[0], obs[0])
    assert ei.value.code == -4
    eng.nprob, eng.nobs = 1, 2  # (the Python mirror only records a successful build)
    assert eng.build_status()[0] != 0
    g, jac = eng.eval(np.zeros(7))
    assert np.all(np.isfinite(jac)) and np.all(g[0][:7 * 128 + 7 * 128 * 2] >= 1e299)
    assert eng.finalize_solution(g[0]) == (False, 0)
    k, ok, first, _ = eng.solve(q_des)
    assert not ok[0]


def test_capacities_must_be_multiples_of_eight(built):
    from armour_b200 import ArmourError, ReachSetEngine
    with pytest.raises(ArmourError):
        ReachSetEngine(max_problems=1, max_obstacles=2, cap_link=2, cap_torque=4)
    with pytest.raises(ArmourError):
        ReachSetEngine(max_problems=70000, max_obstacles=2)
