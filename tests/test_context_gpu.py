"""Context life-cycle hazards of the C ABI (round-1 advisor findings), on the GPU:
* the device solver's scratch follows the obstacle count of the batch (rebuild with more obstacles, solve again);
* two live contexts of DIFFERENT configuration on one device never see each other's constants;
* growing a reservation drops the built batch instead of leaving half-space lists uninitialised."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K_TEST = np.array([0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7])


def test_solver_scratch_follows_the_obstacle_count(built):
    from armour_b200 import ReachSetEngine, worlds
    n = 6
    qa = worlds.random_problems(n, 2, seed=91)
    qb = worlds.random_problems(n, 10, seed=92)
    eng = ReachSetEngine(max_problems=n, max_obstacles=10)
    eng.build(*qa[:3], qa[4])
    eng.solve(qa[3])
    eng.build(*qb[:3], qb[4])  # same context, m grows from 2 716 to 9 884 rows
    k, ok, first, iters = eng.solve(qb[3])
    fresh = ReachSetEngine(max_problems=n, max_obstacles=10)
    fresh.build(*qb[:3], qb[4])
    k2, ok2, first2, iters2 = fresh.solve(qb[3])
    assert np.array_equal(k, k2) and np.array_equal(ok, ok2) and np.array_equal(first, first2) and np.array_equal(iters, iters2)


def test_two_contexts_of_different_configuration_interleaved(built):
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, _, obs = worlds.random_problems(1, 4, seed=93)
    cfg_a = dict()                                                                     # 7 links, default threshold
    cfg_b = dict(robot_model=1, mass_uncertainty=0.10, inertia_uncertainty=0.10)       # gripper link, 10 % payload
    a = ReachSetEngine(max_problems=1, max_obstacles=4, **cfg_a)
    b = ReachSetEngine(max_problems=1, max_obstacles=4, cap_link=64, cap_torque=128, **cfg_b)
    ref_a = OracleProblem().build(q0[0], qd0[0], qdd0[0], obs[0])
    ref_b = OracleProblem(model_id=1, mass_uncertainty=0.10, inertia_uncertainty=0.10).build(q0[0], qd0[0], qdd0[0], obs[0])
    # build and evaluate in alternation: each launch must run on its own context's constants
    a.build(q0[0], qd0[0], qdd0[0], obs[0])
    b.build(q0[0], qd0[0], qdd0[0], obs[0])
    for _ in range(2):
        ga, Ja = a.eval(K_TEST)
        gb, Jb = b.eval(K_TEST)
        assert np.max(np.abs(ga[0] - ref_a.eval_g(K_TEST))) <= 1e-9 and np.max(np.abs(Ja[0] - ref_a.eval_jac_g(K_TEST))) <= 1e-9
        assert np.max(np.abs(gb[0] - ref_b.eval_g(K_TEST))) <= 1e-9 and np.max(np.abs(Jb[0] - ref_b.eval_jac_g(K_TEST))) <= 1e-9
    a.build(q0[0], qd0[0], qdd0[0], obs[0])  # b owned the constants last
    tr, tr_ref = a.torque_radius()[0], ref_a.torque_radius()
    assert np.all(tr >= tr_ref) and np.max((tr - tr_ref) / tr_ref) <= 1e-10


def test_growing_a_reservation_drops_the_built_batch(built):
    from armour_b200 import ArmourError, ReachSetEngine, worlds
    q0, qd0, qdd0, _, obs = worlds.random_problems(2, 3, seed=94)
    eng = ReachSetEngine(max_problems=8, max_obstacles=20)
    eng.build(q0, qd0, qdd0, obs)
    g0, _ = eng.eval(np.zeros((2, 7)))
    rc = eng.lib.armour_ctx_reserve(eng._h, 8, 20)
    assert rc == 0
    with pytest.raises(ArmourError):  # ARMOUR_ERR_STATE: the half-space lists went with the old buffers
        eng.eval(np.zeros((2, 7)))
    eng.build(q0, qd0, qdd0, obs)
    g1, _ = eng.eval(np.zeros((2, 7)))
    assert np.array_equal(g0, g1)
