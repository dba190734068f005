"""ARMTD comparison planner, CPU side: the oracle's restatement (oracle/armtd.cpp) against outputs of the
REFERENCE's own KPA sources compiled by nvcc and run on a B200, frozen in tests/golden/armtd/reference.npz
(tools/make_golden_armtd.py).  Tables, bounds, cost and verdict exactly; collision rows to 1e-14 (the reference's collision
kernels contract into FMAs, nvcc's default, the oracle does not)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "armtd", "reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_oracle_matches_the_frozen_reference_outputs(built, gold, case):
    from oracle.pyoracle import OracleArmtd
    t = f"c{case}_"
    o = OracleArmtd().build(gold[t + "q0"], gold[t + "qd0"], gold[t + "jrs"], gold[t + "k_range"], gold[t + "obs"])
    assert o.m == gold[t + "g"].shape[1] == 7 * 100 * o.nobs + 28
    gl, gu = o.bounds()
    assert np.array_equal(gl, gold[t + "gl"]) and np.array_equal(gu, gold[t + "gu"])
    lg = o.link_gens().transpose(0, 1, 3, 2).reshape(o.T, o.NJ, 18)
    assert np.array_equal(lg, gold[t + "link_gens"])
    n, c, h, g = o.link_tables()
    nm = gold[t + "tab_key"].shape[1]
    assert np.array_equal(n, gold[t + "tab_n"]) and np.array_equal(c, gold[t + "tab_center"])
    assert np.array_equal(h[:, :nm], gold[t + "tab_key"]) and np.array_equal(g[:, :nm], gold[t + "tab_coeff"])
    for i, k in enumerate(gold[t + "k"]):
        gg, J = o.eval_g(k), o.eval_jac_g(k)
        assert np.max(np.abs(gg - gold[t + "g"][i])) <= 1e-14 and np.max(np.abs(J - gold[t + "J"][i])) <= 1e-14
        assert np.array_equal(gg[-28:], gold[t + "g"][i][-28:]) and np.array_equal(J[-28:], gold[t + "J"][i][-28:])
        assert o.verdict(gg)[0] == bool(gold[t + "feasible"][i])
        assert o.cost(gold[t + "q_des"], k) == gold[t + "f"][i]
        assert np.array_equal(o.cost_grad(gold[t + "q_des"], k), gold[t + "df"][i])


def test_fixture_has_feasible_and_infeasible_plans(gold):
    assert gold["c0_feasible"].all() and not gold["c3_feasible"].any()
    assert gold["c1_g"].shape[1] != gold["c0_g"].shape[1]  # worlds with different obstacle counts
