"""GPU parity of the whole product path (K1 build + K3a + K3 through the C ABI) against the REFERENCE ITSELF:
tests/golden/refcuda/*.npz hold g, Jacobian rows, bounds, cost and verdicts computed on a B200 by the reference's own
sources, CUDA collision kernels included (see tests/test_refcuda_golden.py, tools/make_golden_collision.py).

Bars (BASELINE north_star): g and Jacobian within 1e-9, verdict identical, every torque bound interval contained in the
reference's and within 1e-10 relative (the GPU radii are rounded outward).  When oracle/_ref/libarmour_ref_cuda.so
travelled to the box the same comparison is also made live on a problem that has no committed fixture.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
NF, T, NJ = 7, 128, 7
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refcuda", "*.npz")))
TOL = 1e-9


def check_against_reference(eng, q_des, ks, g_of, jac_of, rows, feasible_of, f_of, grad_of, g_l, g_u):
    m = g_l.shape[0]
    c0 = NF * T
    gl, gu = eng.get_bounds_info()
    # torque rows: [-lim + r, lim - r]; GPU r >= reference r (outward rounding), within 1e-10 relative
    assert np.all(gl[0][:c0] >= g_l[:c0]) and np.all(gu[0][:c0] <= g_u[:c0]), "torque bounds must be the tighter ones"
    assert np.max(np.abs(gl[0][:c0] - g_l[:c0]) / np.abs(g_l[:c0])) <= 1e-10
    assert np.max(np.abs(gu[0][:c0] - g_u[:c0]) / np.abs(g_u[:c0])) <= 1e-10
    assert np.array_equal(gl[0][c0:], g_l[c0:]) and np.array_equal(gu[0][c0:], g_u[c0:])
    worst_g = worst_j = 0.0
    for n, k in enumerate(ks):
        g, J = eng.eval(k)
        worst_g = max(worst_g, float(np.max(np.abs(g[0] - g_of(n)))))
        worst_j = max(worst_j, float(np.max(np.abs(J[0][rows] - jac_of(n)))))
        ok, _ = eng.finalize_solution(g[0])
        assert ok == feasible_of(n), f"verdict differs from the reference at k #{n}"
        obj, grad = eng.cost(q_des, k)
        assert abs(obj - f_of(n)) <= 1e-12 and np.max(np.abs(grad - grad_of(n))) <= 1e-12
    assert worst_g <= TOL and worst_j <= TOL, (worst_g, worst_j)
    assert m == eng.m


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_product_matches_reference_planner(built, path):
    from armour_b200 import ReachSetEngine
    gold = dict(np.load(path))
    obs = gold["obstacles"].reshape(-1, 12)
    eng = ReachSetEngine(max_problems=1, max_obstacles=max(obs.shape[0], 1))
    eng.build(gold["q0"], gold["qd0"], gold["qdd0"], obs)
    check_against_reference(eng, gold["q_des"], gold["ks"], lambda n: gold[f"g_{n}"], lambda n: gold[f"jac_{n}"],
                            gold["jac_rows"], lambda n: bool(gold[f"feasible_{n}"]), lambda n: float(gold[f"f_{n}"]),
                            lambda n: gold[f"grad_f_{n}"], gold["g_l"], gold["g_u"])
    eng.close()


def test_product_matches_reference_planner_live(built):
    """Same bars against the reference library running next to the product on this GPU (no fixture in between)."""
    from oracle import pyrefcuda
    if not pyrefcuda.available():
        pytest.skip("oracle/_ref/libarmour_ref_cuda.so did not travel to this box")
    from armour_b200 import ReachSetEngine, worlds
    q0, qd0, qdd0, q_des, obs = worlds.random_problems(3, 10, seed=777)
    ks = np.vstack([np.zeros(7), worlds.halton_k(3, skip=11)])
    for p in range(3):
        ref = pyrefcuda.ReferencePlanner(q0[p], qd0[p], qdd0[p], q_des[p], obs[p])
        eng = ReachSetEngine(max_problems=1, max_obstacles=10)
        eng.build(q0[p], qd0[p], qdd0[p], obs[p])
        gs = [ref.eval_g(k) for k in ks]
        Js = [ref.eval_jac_g(k) for k in ks]
        fs = [ref.cost(k) for k in ks]
        _, _, g_l, g_u = ref.bounds()
        check_against_reference(eng, q_des[p], ks, lambda n: gs[n], lambda n: Js[n], np.arange(ref.m),
                                lambda n: ref.finalize(ks[n], gs[n]), lambda n: fs[n][0], lambda n: fs[n][1], g_l, g_u)
        eng.close()
