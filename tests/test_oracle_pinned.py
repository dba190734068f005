"""Pins the restated oracle (oracle/*.cpp) against the REFERENCE's own implementation.

* tests/golden/ref/*.npz were produced by tools/make_golden.py from oracle/_ref: the reference's PZsparse.cu,
  Trajectory.cu and Dynamics.cu compiled from /root/reference against stand-in Eigen / Boost.Interval headers.
  The oracle must reproduce them BIT FOR BIT (tables, radii, torque radius, host-side rows of eval_g /
  eval_jac_g); only the generated Bezier-extremum derivative expressions may differ by rounding (1e-12).
* when oracle/_ref is present (this container; it also travels to the GPU box) the same comparison is made
  live on a problem that has no committed fixture.
"""
import os

import numpy as np
import pytest

import golden_util as gu
from conftest import WORLDS

NF = 7


@pytest.mark.parametrize("path", gu.FIXTURES, ids=[os.path.basename(p)[:-4] for p in gu.FIXTURES])
def test_oracle_reproduces_reference_fixture(built, path):
    from oracle.pyoracle import OracleProblem
    gold = gu.load(path)
    orc = OracleProblem().build(gold["q0"], gold["qd0"], gold["qdd0"], np.zeros((0, 12)))
    gu.check_tables(orc.tables(64, 128), gold, exact=True)
    ts = gold["t_subset"]
    sel_u = np.array([t * NF + j for t in ts for j in range(NF)])
    for n, k in enumerate(gold["ks"]):
        g, J = orc.eval_g(k), orc.eval_jac_g(k)
        assert np.array_equal(g[sel_u], gold[f"g_torque_{n}"])
        assert np.array_equal(J[sel_u], gold[f"jac_torque_{n}"])
        assert np.array_equal(orc.link_sliced_center()[ts], gold[f"link_c_{n}"])
        assert np.array_equal(g[-28:], gold[f"bez_{n}"])
        assert np.max(np.abs(J[-28:] - gold[f"dbez_{n}"])) <= 1e-12


def test_fixture_set_is_complete():
    assert len(gu.FIXTURES) >= 4


def test_oracle_matches_live_reference_build(built):
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    from armour_b200 import worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, _, obs = worlds.random_problems(1, 4, seed=424242)
    ref = pyref.ReferenceProblem(q0[0], qd0[0], qdd0[0])
    orc = OracleProblem().build(q0[0], qd0[0], qdd0[0], obs[0])
    R, O = ref.tables(), orc.tables()
    for key in R:
        assert np.array_equal(R[key], O[key]), key
    k = worlds.halton_k(1, skip=5)[0]
    s, g, J = ref.slice(k), orc.eval_g(k), orc.eval_jac_g(k)
    assert np.array_equal(s["g_torque"], g[:896]) and np.array_equal(s["jac_torque"], J[:896])
    assert np.array_equal(s["link_c"], orc.link_sliced_center())
    assert np.array_equal(s["bez"], g[-28:]) and np.max(np.abs(s["dbez"] - J[-28:])) <= 1e-12
