"""ARMTD comparison planner on the GPU through the C ABI (armour_armtd_*): against the frozen outputs of the
reference's own KPA sources (tests/golden/armtd/reference.npz) and the oracle on a fresh problem."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "armtd", "reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_planner_matches_the_frozen_reference_outputs(built, gold, case):
    from armour_b200 import ArmtdPlanner
    t = f"c{case}_"
    p = ArmtdPlanner().build(gold[t + "q0"], gold[t + "qd0"], gold[t + "jrs"], gold[t + "k_range"], gold[t + "obs"])
    assert p.m == gold[t + "g"].shape[1]
    gl, gu = p.get_bounds_info()
    assert np.array_equal(gl, gold[t + "gl"]) and np.array_equal(gu, gold[t + "gu"])
    # reduce_link_PZ output: generators exactly, radii within the outward rounding of the build kernel (2^-40 relative)
    lg = p.link_independent_generators()
    ref = gold[t + "link_gens"]
    assert np.max(np.abs(lg - ref) / (1e-300 + np.abs(ref) + 1e-12)) <= 1e-9 and np.all(lg[..., 9:] >= ref[..., 9:] - 1e-15)
    for i, k in enumerate(gold[t + "k"]):
        g, J = p.eval(k)
        assert np.max(np.abs(g - gold[t + "g"][i])) <= 1e-11, f"k {i}"
        assert np.max(np.abs(J - gold[t + "J"][i])) <= 1e-11, f"k {i}"
        assert np.array_equal(g[-28:], gold[t + "g"][i][-28:]) and np.array_equal(J[-28:], gold[t + "J"][i][-28:])
        assert np.max(np.abs(p.link_sliced_center() - gold[t + "link_sliced_center"][i])) <= 1e-12
        assert np.array_equal(p.eval_g(k), g) and np.array_equal(p.eval_jac_g(k), J)
        assert p.finalize_solution(g)[0] == bool(gold[t + "feasible"][i])
        assert p.eval_f(gold[t + "q_des"], k) == gold[t + "f"][i]
        assert np.array_equal(p.eval_grad_f(gold[t + "q_des"], k), gold[t + "df"][i])
    p.close()


def test_fresh_problem_against_the_oracle_and_verdict_rows(built):
    from armour_b200 import ArmtdPlanner, worlds
    from oracle.pyoracle import OracleArmtd
    q0, qd0, q_des, jrs, kr, obs = worlds.armtd_problem(os.path.join(HERE, "golden", "worlds", "scene_031_004.csv"), seed=9)
    p = ArmtdPlanner().build(q0, qd0, jrs, kr, obs)
    o = OracleArmtd().build(q0, qd0, jrs, kr, obs)
    assert p.m == o.m
    rng = np.random.default_rng(2)
    for k in np.vstack([np.zeros(7), rng.uniform(-1, 1, (4, 7))]):
        g, J = p.eval(k)
        go, Jo = o.eval_g(k), o.eval_jac_g(k)
        assert np.max(np.abs(g - go)) <= 1e-11 and np.max(np.abs(J - Jo)) <= 1e-11
        assert p.finalize_solution(go) == o.verdict(go)
        # violated rows of every class give the same first-violation index as the oracle's verdict
        for row in (3, 6 * 100 * o.nobs + 5, o.m - 28 + 2, o.m - 14 + 1, o.m - 7 + 6):
            gv = go.copy()
            gv[row] = 1e3
            assert p.finalize_solution(gv) == o.verdict(gv)
    # a second build in the same context (other world, other obstacle count)
    q0, qd0, q_des, jrs, kr, obs = worlds.armtd_problem(os.path.join(HERE, "golden", "worlds", "scene_013_001.csv"), seed=3)
    p.build(q0, qd0, jrs, kr, obs)
    o = OracleArmtd().build(q0, qd0, jrs, kr, obs)
    g, J = p.eval(np.full(7, 0.3))
    assert np.max(np.abs(g - o.eval_g(np.full(7, 0.3)))) <= 1e-11
    p.close()


def test_argument_errors(built):
    from armour_b200 import ArmourError, ArmtdPlanner, ReachSetEngine
    p = ArmtdPlanner(max_obstacles=2)
    with pytest.raises(ValueError):
        p.build(np.zeros(7), np.zeros(7), np.zeros((6, 7, 99)), np.ones(7), np.zeros((1, 12)))
    with pytest.raises(ArmourError):
        p.build(np.zeros(7), np.zeros(7), np.zeros((6, 7, 100)), np.ones(7), np.zeros((3, 12)))  # more than max_obstacles
    p.close()
    eng = ReachSetEngine(max_problems=1, max_obstacles=2)
    assert eng.lib.armour_armtd_num_constraints(eng._h) < 0  # not an ARMTD context
    eng.close()


def test_armtd_main_cli_matches_the_cpu_planner(built, tmp_path):
    """armtd_main (the drop-in for KPA/armtd_main.cu): armtd.in -> the four output files, against the same local solver
    driving the oracle on the CPU."""
    import subprocess

    from armour_b200 import worlds
    from oracle.pyoracle import OracleArmtd
    exe = os.path.join(os.path.dirname(HERE), "armour_b200", "armtd_main")
    q0, qd0, q_des, jrs, kr, obs = worlds.armtd_problem(os.path.join(HERE, "golden", "worlds", "scene_016_006.csv"), seed=1)
    worlds.write_armtd_in(str(tmp_path / "armtd.in"), q0, qd0, q_des, jrs, kr, obs)
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    o = OracleArmtd().build(q0, qd0, jrs, kr, obs)
    k_ref, ok_ref, _, _ = o.solve(q_des)
    lines = (tmp_path / "armtd.out").read_text().split()
    if "wall time exceeded" in res.stdout:
        pytest.skip("the cold process ran into the reference's 0.4 s optimiser budget")
    assert ok_ref and len(lines) == 8  # 7 x k_opt + the time
    k = np.array([float(x) for x in lines[:7]])
    assert np.max(np.abs(k - k_ref)) <= 1e-8
    cen = np.loadtxt(tmp_path / "armtd_joint_position_center.out").reshape(100, 7, 3)
    assert np.max(np.abs(cen - o.link_sliced_center())) <= 1e-8
    rad = np.loadtxt(tmp_path / "armtd_joint_position_radius.out").reshape(100, 7, 3, 6)
    ref = o.link_gens()  # [T, NJ, 3, 6]
    assert np.max(np.abs(rad - ref)) <= 1e-8
    g = np.loadtxt(tmp_path / "armtd_constraints.out")
    assert g.shape == (o.m,) and np.max(np.abs(g - o.eval_g(k_ref))) <= 1e-5
    # too many obstacles: a single -1, exit code -1 (255)
    worlds.write_armtd_in(str(tmp_path / "armtd.in"), q0, qd0, q_des, jrs, kr, np.zeros((41, 12)))
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode != 0 and (tmp_path / "armtd.out").read_text().strip() == "-1"
