"""BASELINE config 2 at FULL size (1,024 random worlds x 10 obstacles, the bench batch) through size-independent
properties: no oracle run is affordable at this size, so the CUDA path is checked against itself and against the
mathematics of the constraints — derivative consistency, verdict = bounds, batch independence, export/import round
trip, run-to-run determinism."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, NOBS = 1024, 10


@pytest.fixture(scope="module")
def batch(built):
    from armour_b200 import ReachSetEngine, worlds
    q0, qd0, qdd0, q_des, obs = worlds.random_problems(N, NOBS, seed=20261017)
    eng = ReachSetEngine(max_problems=N, max_obstacles=NOBS)
    eng.build(q0, qd0, qdd0, obs)
    assert not eng.build_status().any(), "capacity failure in the bench batch"
    return eng, (q0, qd0, qdd0, q_des, obs)


def test_build_is_deterministic_and_sane(batch):
    eng, (q0, qd0, qdd0, q_des, obs) = batch
    tr = eng.torque_radius().copy()
    ln, un = eng.monomial_counts()
    cnt = eng.candidate_counts().copy()
    assert np.all(np.isfinite(tr)) and np.all(tr > 0)
    assert ln.min() >= 0 and ln.max() <= 32 and un.min() >= 1 and un.max() <= 64
    assert cnt.min() >= 1 and (cnt == 255).sum() == 0 and cnt.mean() < 4.0
    eng.build(q0, qd0, qdd0, obs)  # run to run: bit-identical
    assert np.array_equal(eng.torque_radius(), tr)
    ln2, un2 = eng.monomial_counts()
    assert np.array_equal(ln2, ln) and np.array_equal(un2, un) and np.array_equal(eng.candidate_counts(), cnt)


def test_jacobian_is_the_derivative_of_g(batch):
    """Central differences of g (14 extra evaluations of all 1,024 worlds) against the analytic Jacobian.  Torque rows
    are polynomials in k; a collision row is the negated maximum of a few half-space values, smooth except where the
    winning half-space changes, so a small share of those rows may sit on a kink."""
    from armour_b200 import worlds
    eng, _ = batch
    k = 0.8 * worlds.halton_k(N, skip=7)
    g0, jac = eng.eval(k)
    T, m = eng.T, eng.m
    h = 1e-5
    n_tq = 7 * T
    bad_collision = bad_limits = 0
    for j in range(7):
        e = np.zeros(7)
        e[j] = h
        gp, _ = eng.eval(k + e, True, False)
        gm, _ = eng.eval(k - e, True, False)
        fd = (gp - gm) / (2 * h)
        err = np.abs(fd - jac[:, :, j])
        scale = 1.0 + np.abs(jac[:, :, j])
        assert np.max(err[:, :n_tq] / scale[:, :n_tq]) < 1e-6, f"torque rows, variable {j}"
        # (the joint-limit rows are extrema over the trajectory: piecewise smooth too, where the extremal point switches)
        bad_limits += int((err[:, m - 28:] / scale[:, m - 28:] > 1e-5).sum())
        bad_collision += int((err[:, n_tq:m - 28] / scale[:, n_tq:m - 28] > 1e-5).sum())
    assert bad_collision < 1e-3 * 7 * N * (m - 28 - n_tq), bad_collision
    assert bad_limits < 1e-2 * 7 * N * 28, bad_limits


def test_verdict_equals_the_bounds(batch):
    """finalize_solution's predicate (device kernel) = g within [g_l - tol, g_u + tol] row class by row class, and the
    first violated row is the smallest such index (KPR/NLPclass.cu:449-537)."""
    import torch
    from armour_b200 import worlds
    eng, _ = batch
    k = worlds.halton_k(N, skip=3)
    g, _ = eng.eval(k, True, False)
    gl, gu = eng.get_bounds_info()
    T, m = eng.T, eng.m
    tol = np.zeros(m)
    tol[:7 * T] = 1e-2            # TORQUE_INPUT_VIOLATION_THRESHOLD
    tol[7 * T:m - 28] = 1e-4      # COLLISION_AVOIDANCE_VIOLATION_THRESHOLD
    bad = (g > gu + tol) | (g < gl - tol)
    ok_ref = ~bad.any(axis=1)
    first_ref = np.where(ok_ref, -1, bad.argmax(axis=1))
    d_g = torch.from_numpy(g).cuda()
    d_ok = torch.empty(N, dtype=torch.int32, device="cuda")
    d_first = torch.empty(N, dtype=torch.int32, device="cuda")
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.verdict_device(N, d_g.data_ptr(), d_ok.data_ptr(), d_first.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_ok.cpu().numpy().astype(bool), ok_ref)
    assert np.array_equal(d_first.cpu().numpy(), first_ref)
    assert 0 < ok_ref.sum() < N  # the batch has both kinds


def test_problems_do_not_see_each_other(batch):
    """Rows of a problem do not depend on what else is in the launch: the first 16 worlds evaluated alone give the
    bits of the full launch; an exported reach set imported into a fresh context evaluates to the same bits."""
    from armour_b200 import ReachSetEngine, worlds
    eng, (q0, qd0, qdd0, q_des, obs) = batch
    k = worlds.halton_k(N, skip=19)
    g, jac = eng.eval(k)
    g16, jac16 = eng.eval(k[:16])
    assert np.array_equal(g16, g[:16]) and np.array_equal(jac16, jac[:16])
    for p in (0, 511, N - 1):
        tables = eng.export_reachsets(p)
        one = ReachSetEngine(max_problems=1, max_obstacles=NOBS)
        one.import_reachsets(0, 1, tables, q0[p], qd0[p], qdd0[p], obs[p])
        g1, j1 = one.eval(k[p])
        assert np.array_equal(g1[0], g[p]) and np.array_equal(j1[0], jac[p])


def test_sampled_worlds_of_the_bench_batch_match_the_oracle(batch):
    """Eight worlds of THE batch bench.py times (seed 20261017), spread over it, against the CPU oracle: g, Jacobian and
    verdict of the batched build + batched evaluation at two k per world."""
    from armour_b200 import worlds
    from oracle.pyoracle import OracleProblem
    eng, (q0, qd0, qdd0, q_des, obs) = batch
    ks = worlds.halton_k(2 * N, skip=7).reshape(2, N, 7)
    ks[0] *= 0.0
    sample = [0, 1, 127, 300, 511, 640, 900, 1023]
    for it in range(2):
        g, jac = eng.eval(ks[it])
        for p in sample:
            ref = OracleProblem().build(q0[p], qd0[p], qdd0[p], obs[p])
            g_ref, j_ref = ref.eval_g(ks[it, p]), ref.eval_jac_g(ks[it, p])
            assert np.max(np.abs(g[p] - g_ref)) <= 1e-9, (it, p)
            assert np.max(np.abs(jac[p] - j_ref)) <= 1e-9, (it, p)
            # the verdict of problem p alone (armour_verdict works on problem 0 of a context: use the bounds)
            gl, gu = ref.bounds()
            ok_ref, first_ref = ref.verdict(g_ref)
            ok_gpu, first_gpu = ref.verdict(g[p])
            assert (ok_ref, first_ref) == (ok_gpu, first_gpu)
