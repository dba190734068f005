"""Robust-controller path, CPU side: the oracle's restatement (oracle/controller.cpp) against outputs of the
REFERENCE's own sources — frozen in tests/golden/controller/reference.npz (tools/make_golden_controller.py) and, where
oracle/_ref/libarmour_ref_controller.so exists, live on fresh random states.  Interval end points must agree bit for bit."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "controller", "reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("eps", [0.03, 0.0, 0.1])
def test_oracle_matches_the_frozen_reference_outputs(built, gold, eps):
    from oracle.pycontroller import OracleController
    o = OracleController(eps=eps)
    tag = f"eps{eps}"
    assert np.array_equal(o.interval_model(), gold[tag + "_model"])
    n = gold["q"].shape[0]
    for i in range(n):
        a = [gold[k][i] for k in ("q", "qd", "qda", "qdd")]
        assert np.array_equal(o.rnea(*a), gold[tag + "_tau"][i])
        lo, hi = o.rnea_interval(*a)
        assert np.array_equal(lo, gold[tag + "_lo"][i]) and np.array_equal(hi, gold[tag + "_hi"][i])
        lo, hi = o.rnea_interval(*a, gravity=False)
        assert np.array_equal(lo, gold[tag + "_lo_nograv"][i]) and np.array_equal(hi, gold[tag + "_hi_nograv"][i])
        for name in ("a", "b"):
            alpha, V_max, thr = gold[f"{tag}_{name}_gains"]
            u, un, v, st = o.update(gold["Kr"], alpha, V_max, thr, gold["q"][i], gold["qd"][i], gold["q_des"][i], gold["qd_des"][i],
                                    gold["qdd_des"][i])
            assert np.array_equal(u, gold[f"{tag}_{name}_u"][i]) and np.array_equal(un, gold[f"{tag}_{name}_un"][i])
            assert np.array_equal(v, gold[f"{tag}_{name}_v"][i]) and st == gold[f"{tag}_{name}_status"][i]


def test_fixture_covers_the_branches(gold):
    """the frozen cases include states below the |r| threshold (v = 0), active robust inputs and wrapped position errors"""
    v = gold["eps0.03_a_v"]
    assert np.any(np.all(v == 0, axis=1)) and np.any(np.any(v != 0, axis=1))
    assert np.any(np.abs(gold["q_des"] - gold["q"]) > np.pi)
    assert np.all(gold["eps0.03_lo"] <= gold["eps0.03_tau"]) and np.all(gold["eps0.03_tau"] <= gold["eps0.03_hi"])
    assert np.array_equal(gold["eps0.0_lo"] <= gold["eps0.0_tau"], np.ones_like(gold["eps0.0_tau"], dtype=bool))


def test_oracle_matches_the_reference_library_live(built):
    from oracle import pycontroller
    if not pycontroller.reference_available():
        pytest.skip("oracle/_ref/libarmour_ref_controller.so not built (needs /root/reference at build time)")
    o, r = pycontroller.OracleController(), pycontroller.ReferenceController()
    rng = np.random.default_rng(5)
    for t in range(100):
        q = rng.uniform(-np.pi, np.pi, 7)
        qd, qda, qdd = (rng.uniform(-2, 2, 7) for _ in range(3))
        assert np.array_equal(o.rnea(q, qd, qda, qdd), r.rnea(q, qd, qda, qdd))
        for a, b in zip(o.rnea_interval(q, qd, qda, qdd), r.rnea_interval(q, qd, qda, qdd)):
            assert np.array_equal(a, b)
        q_des, qd_des, qdd_des = q + rng.uniform(-0.1, 0.1, 7), qd + rng.uniform(-0.1, 0.1, 7), rng.uniform(-2, 2, 7)
        Kr = rng.uniform(1, 20, 7)
        a = o.update(Kr, 1.0, 1e-2, 1e-10, q, qd, q_des, qd_des, qdd_des)
        b = r.update(Kr, 1.0, 1e-2, 1e-10, q, qd, q_des, qd_des, qdd_des)
        assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3]


def test_interval_torque_contains_sampled_models(built):
    """what the interval pass is for: the torque of ANY model with masses / inertias inside +-eps lies in the interval torque.
    Checked with the nominal pass of models whose file masses and inertias are scaled by factors in [1 - eps, 1 + eps]."""
    import tempfile

    from oracle.pycontroller import MODEL, OracleController
    eps = 0.03
    o = OracleController(eps=eps)
    rng = np.random.default_rng(11)
    text = open(MODEL).read().splitlines()
    states = [(rng.uniform(-np.pi, np.pi, 7),) + tuple(rng.uniform(-1.5, 1.5, 7) for _ in range(3)) for _ in range(6)]
    bounds = [o.rnea_interval(*s) for s in states]
    for trial in range(6):
        lines = []
        for ln in text:
            if ln.startswith("inertia"):
                head, body = ln.split("<")
                vals = [float(x) for x in body.rstrip(">").split()]
                # the file holds the inertia about the JOINT frame; the reference scales the converted (CoM-frame) entries, so
                # scale the whole rigid body: mass, inertia and m*c_hat by one factor keeps the CoM and scales every converted
                # entry by that factor
                f = 1 + eps * rng.uniform(-1, 1)
                vals = [v * f for v in vals]
                ln = head + "<" + " ".join(repr(v) for v in vals) + ">"
            lines.append(ln)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
            fh.write("\n".join(lines) + "\n")
        try:
            m = OracleController(model_file=fh.name, eps=0.0)
            for s, (lo, hi) in zip(states, bounds):
                tau = m.rnea(*s)
                assert np.all(tau >= lo - 1e-12) and np.all(tau <= hi + 1e-12)
        finally:
            os.unlink(fh.name)
