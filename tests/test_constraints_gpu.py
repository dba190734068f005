"""GPU parity of the constraint kernels (K3a hyper-planes + K3 slice-and-evaluate) through the C ABI.

Reach sets are built by the CPU oracle and uploaded with armour_import_reachsets, so these tests pin the
constraint side in isolation: g(k), the dense Jacobian, the sliced link centres, bounds, verdict and
cost must agree with the oracle.  Tolerance: 1e-12 absolute (the kernels follow the oracle's operation
order without FMA contraction; the only freedom is k^3 and the Bezier powers, a few ulp).
BASELINE tolerance for these quantities is 1e-9.
"""
import glob
import os

import numpy as np
import pytest

from conftest import WORLDS

pytestmark = pytest.mark.gpu
TOL = 1e-12
K_TEST = np.array([0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7])  # reference PZ_tests.cu:198


def _problems():
    from armour_b200 import worlds
    probs = [worlds.config1_problem(os.path.join(WORLDS, "scene_016_006.csv")),
             worlds.config1_problem(os.path.join(WORLDS, "scene_013_001.csv"))]
    q0, qd0, qdd0, qdes, obs = worlds.random_problems(2, 10, seed=5)
    for p in range(2):
        probs.append((q0[p], qd0[p], qdd0[p], qdes[p], obs[p]))
    return probs


@pytest.mark.parametrize("pi", range(4))
def test_import_eval_matches_oracle(built, pi):
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, q_des, obs = _problems()[pi]
    ref = OracleProblem().build(q0, qd0, qdd0, obs)
    eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
    eng.import_reachsets(0, 1, ref.tables(), q0, qd0, qdd0, obs)
    assert eng.m == ref.m
    ks = np.vstack([np.zeros(7), K_TEST, worlds.halton_k(4), np.ones(7), -np.ones(7)])
    for k in ks:
        g, jac = eng.eval(k)
        g_ref, j_ref = ref.eval_g(k), ref.eval_jac_g(k)
        assert np.max(np.abs(g[0] - g_ref)) <= TOL
        assert np.max(np.abs(jac[0] - j_ref)) <= TOL
        assert np.max(np.abs(eng.link_sliced_center() - ref.link_sliced_center())) <= TOL
        assert eng.finalize_solution(g[0]) == ref.verdict(g_ref)
        # separate entry points agree with the fused one
        assert np.array_equal(eng.eval_g(k)[0], g[0])
        assert np.array_equal(eng.eval_jac_g(k)[0], jac[0])
        obj, grad = eng.cost(q_des, k)
        assert abs(obj - ref.cost(q_des, k)) <= 1e-12
        assert np.max(np.abs(grad - ref.cost_grad(q_des, k))) <= 1e-12
    gl, gu = eng.get_bounds_info()
    gl_ref, gu_ref = ref.bounds()
    assert np.array_equal(gl[0], gl_ref) and np.array_equal(gu[0], gu_ref)


def test_batched_import_eval(built):
    """Three problems in one context, a different k per problem, device-pointer path with torch tensors."""
    import torch
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    probs = _problems()[:3]
    # same obstacle count within a batch: trim to the smallest
    nobs = min(p[4].shape[0] for p in probs)
    refs = []
    eng = ReachSetEngine(max_problems=3, max_obstacles=nobs)
    for i, (q0, qd0, qdd0, q_des, obs) in enumerate(probs):
        obs = obs[:nobs]
        ref = OracleProblem().build(q0, qd0, qdd0, obs)
        refs.append(ref)
        eng.import_reachsets(i, 3, ref.tables(), q0, qd0, qdd0, obs)
    ks = worlds.halton_k(3, skip=11)
    m = eng.m
    d_k = torch.tensor(ks, dtype=torch.float64, device="cuda")
    d_g = torch.empty((3, m), dtype=torch.float64, device="cuda")
    d_j = torch.empty((3, m, 7), dtype=torch.float64, device="cuda")
    d_ok = torch.empty(3, dtype=torch.int32, device="cuda")
    d_first = torch.empty(3, dtype=torch.int32, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.eval_device(3, d_k.data_ptr(), d_g.data_ptr(), d_j.data_ptr())
    eng.verdict_device(3, d_g.data_ptr(), d_ok.data_ptr(), d_first.data_ptr())
    torch.cuda.synchronize()
    g, jac = d_g.cpu().numpy(), d_j.cpu().numpy()
    for i, ref in enumerate(refs):
        g_ref = ref.eval_g(ks[i])
        assert np.max(np.abs(g[i] - g_ref)) <= TOL
        assert np.max(np.abs(jac[i] - ref.eval_jac_g(ks[i]))) <= TOL
        ok_ref, first_ref = ref.verdict(g_ref)
        assert bool(d_ok[i].item()) == ok_ref and int(d_first[i].item()) == first_ref
    # host-buffer path gives the same numbers
    g2, j2 = eng.eval(ks)
    assert np.array_equal(g2, g) and np.array_equal(j2, jac)


def test_k_outside_the_box_uses_the_full_scan(built):
    """The stored half-space candidate lists are exact for |k_j| <= 1 (the NLP's variable bounds,
    NLPclass.cu:86-110).  Outside the box the kernel scans all 72 half-spaces from the generators: same answer
    as the oracle there too, and the two paths agree on the boundary."""
    from armour_b200 import ReachSetEngine
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, q_des, obs = _problems()[2]
    ref = OracleProblem().build(q0, qd0, qdd0, obs)
    eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
    eng.import_reachsets(0, 1, ref.tables(), q0, qd0, qdd0, obs)
    for k in (1.5 * K_TEST, np.full(7, -1.25), np.array([1.0, -1.0, 1.0, -1.0, 1.0, -1.0, 1.0 + 2e-6])):
        g, jac = eng.eval(k)
        assert np.max(np.abs(g[0] - ref.eval_g(k))) <= TOL
        assert np.max(np.abs(jac[0] - ref.eval_jac_g(k))) <= TOL


def test_error_paths(built):
    from armour_b200 import ReachSetEngine, ArmourError
    eng = ReachSetEngine(max_problems=1, max_obstacles=2)
    with pytest.raises(ArmourError) as ei:  # evaluate before build
        eng.eval(np.zeros(7))
    assert ei.value.code == -5
    with pytest.raises(ArmourError) as ei:  # too many obstacles (reference throws, CollisionChecking.cu:10-13)
        eng.build(np.zeros(7), np.zeros(7), np.zeros(7), np.zeros((3, 12)))
    assert ei.value.code == -3


def test_candidate_lists_are_exact_over_the_whole_box(built):
    """The build keeps only the half-spaces that can attain a row's maximum somewhere in the k box (interval
    bound + pairwise test against the best one).  Dropping a half-space that wins anywhere would change g or
    its gradient there: compare with the oracle's full 72-half-space scan on a dense sample of the box, its
    corners and points on its faces."""
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    rng = np.random.default_rng(12)
    for pi in (0, 3):
        q0, qd0, qdd0, q_des, obs = _problems()[pi]
        ref = OracleProblem().build(q0, qd0, qdd0, obs)
        eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
        eng.import_reachsets(0, 1, ref.tables(), q0, qd0, qdd0, obs)
        cnt = eng.candidate_counts()
        assert cnt.max() != 255 and cnt.min() >= 1
        assert cnt.mean() < 6.0, cnt.mean()   # the point of the lists: a few half-spaces per row instead of 72
        corners = rng.choice([-1.0, 1.0], size=(24, 7))
        faces = rng.uniform(-1, 1, size=(24, 7))
        faces[np.arange(24), rng.integers(0, 7, 24)] = rng.choice([-1.0, 1.0], 24)
        ks = np.vstack([worlds.halton_k(48, skip=3), corners, faces])
        for k in ks:
            g, jac = eng.eval(k)
            assert np.max(np.abs(g[0] - ref.eval_g(k))) <= TOL
            assert np.max(np.abs(jac[0] - ref.eval_jac_g(k))) <= TOL


ALL_WORLDS = sorted(glob.glob(os.path.join(WORLDS, "scene_*.csv")))


@pytest.mark.parametrize("path", ALL_WORLDS, ids=[os.path.basename(p)[:-4] for p in ALL_WORLDS])
def test_every_saved_world_on_the_k_schedule(built, path):
    """every saved world copied into the repository (the 13 ten-obstacle worlds plus two others, 6
    to 11 obstacles) x the k schedule (k = 0, the PZ_tests point, Halton points, a corner): the WHOLE product path
    (K1 build + K3a + K3) against the oracle — g, Jacobian, bounds containment, verdict with its first violated row."""
    from armour_b200 import ReachSetEngine, worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, q_des, obs = worlds.config1_problem(path)
    ref = OracleProblem().build(q0, qd0, qdd0, obs)
    eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
    eng.build(q0, qd0, qdd0, obs)
    assert eng.m == ref.m
    ks = np.vstack([np.zeros(7), K_TEST, worlds.halton_k(3, skip=5), [1, -1, 1, 1, -1, 1, -1]])
    for k in ks:
        g, jac = eng.eval(k)
        g_ref, j_ref = ref.eval_g(k), ref.eval_jac_g(k)
        assert np.max(np.abs(g[0] - g_ref)) <= 1e-9 and np.max(np.abs(jac[0] - j_ref)) <= 1e-9
        assert eng.finalize_solution(g[0]) == ref.verdict(g_ref)
    tr, tr_ref = eng.torque_radius()[0], ref.torque_radius()
    assert np.all(tr >= tr_ref) and np.max((tr - tr_ref) / tr_ref) <= 1e-10
    eng.close()


@pytest.mark.parametrize("model", [0, 1])
def test_structured_jacobian_is_the_dense_one_without_its_exact_zeros(built, model):
    """armour_jacobian_structure / armour_batch_eval_structured: the non-zeros returned are the dense entries at (iRow, jCol),
    and every dense entry outside the structure is an exact zero (a collision row of link l depends on k_0..k_l only)."""
    from armour_b200 import ReachSetEngine, worlds
    n, nobs = 3, 6
    q0, qd0, qdd0, _, obs = worlds.random_problems(n, nobs, seed=61)
    eng = ReachSetEngine(max_problems=n, max_obstacles=nobs, robot_model=model, cap_link=64, cap_torque=128)
    eng.build(q0, qd0, qdd0, obs)
    ks = worlds.halton_k(n, skip=9)
    g, J = eng.eval(ks)
    g2, vals = eng.eval_structured(ks)
    ir, jc = eng.jacobian_structure()
    NJ = eng.NJ
    assert eng.jacobian_nnz == 49 * 128 + 128 * nobs * sum(min(l + 1, 7) for l in range(NJ)) + 28 == ir.size
    assert np.array_equal(g, g2)
    mask = np.zeros((eng.m, 7), bool)
    mask[ir, jc] = True
    assert mask.sum() == ir.size, "duplicate entries in the structure"
    for p in range(n):
        assert np.array_equal(vals[p], J[p][ir, jc])
        assert np.all(J[p][~mask] == 0.0)
    # single-problem entry point
    single = np.empty(eng.jacobian_nnz)
    eng._check(eng.lib.armour_eval_jac_g_structured(eng._h, ks[0].ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)),
                                                    single.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double))))
    assert np.array_equal(single, vals[0])
