"""GPU reach sets against the REFERENCE fixtures (tests/golden/ref, generated from the reference's own
sources by tools/make_golden.py) — no oracle in the loop.  Tolerances of BASELINE.json's north_star: every
GPU radius contains the reference's and is within 1e-10 relative; coefficients, centres, g and Jacobian rows
within 1e-9 (asserted at 1e-12); monomial key sets identical."""
import os

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
NF = 7


@pytest.mark.parametrize("path", gu.FIXTURES, ids=[os.path.basename(p)[:-4] for p in gu.FIXTURES])
def test_gpu_build_reproduces_reference_fixture(built, path):
    from armour_b200 import ReachSetEngine
    gold = gu.load(path)
    eng = ReachSetEngine(max_problems=1, max_obstacles=1, cap_link=64, cap_torque=128)
    eng.build(gold["q0"], gold["qd0"], gold["qdd0"], np.zeros((0, 12)))
    gu.check_tables(eng.export_reachsets(0), gold, exact=False, coef_tol=1e-12, rel_radius=1e-10)
    ts = gold["t_subset"]
    sel_u = np.array([t * NF + j for t in ts for j in range(NF)])
    for n, k in enumerate(gold["ks"]):
        g, J = eng.eval(k)
        assert np.max(np.abs(g[0][sel_u] - gold[f"g_torque_{n}"])) <= 1e-12
        assert np.max(np.abs(J[0][sel_u] - gold[f"jac_torque_{n}"])) <= 1e-12
        assert np.max(np.abs(eng.link_sliced_center()[ts] - gold[f"link_c_{n}"])) <= 1e-12
        assert np.max(np.abs(g[0][-28:] - gold[f"bez_{n}"])) <= 1e-12
        assert np.max(np.abs(J[0][-28:] - gold[f"dbez_{n}"])) <= 1e-10
