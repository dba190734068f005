"""Robust-controller path on the GPU through the C ABI: batched passRNEA / passRNEA_Int / RobustController::update
against the reference's frozen outputs (tests/golden/controller/reference.npz) and the oracle on fresh states.
Host-pointer calls (sin / cos from the host's libm, like the reference): interval end points bit for bit.
Device-pointer calls with the device's sincos: relative 1e-12."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
MODEL = os.path.join(HERE, "golden", "robot_models", "kinova_without_gripper.txt")
GOLD = os.path.join(HERE, "golden", "controller", "reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("eps", [0.03, 0.0, 0.1])
def test_batched_calls_match_the_frozen_reference_outputs(built, gold, eps):
    from armour_b200 import RobustController
    c = RobustController(MODEL, eps)
    tag = f"eps{eps}"
    assert np.array_equal(c.interval_model(), gold[tag + "_model"])
    q, qd, qda, qdd = (gold[k] for k in ("q", "qd", "qda", "qdd"))
    tau, lo, hi = c.rnea(q, qd, qda, qdd)
    assert np.array_equal(lo, gold[tag + "_lo"]) and np.array_equal(hi, gold[tag + "_hi"])
    assert np.max(np.abs(tau - gold[tag + "_tau"])) <= 1e-12  # nominal pass: doubles, tolerance (oracle/controller.cpp header)
    _, lo, hi = c.rnea(q, qd, qda, qdd, gravity=False, nominal=False)
    assert np.array_equal(lo, gold[tag + "_lo_nograv"]) and np.array_equal(hi, gold[tag + "_hi_nograv"])
    for name in ("a", "b"):
        alpha, V_max, thr = gold[f"{tag}_{name}_gains"]
        u, un, v, st = c.update(gold["Kr"], alpha, V_max, thr, gold["q"], gold["qd"], gold["q_des"], gold["qd_des"], gold["qdd_des"])
        assert np.array_equal(st, gold[f"{tag}_{name}_status"])
        assert np.max(np.abs(un - gold[f"{tag}_{name}_un"])) <= 1e-12
        scale = 1 + np.abs(gold[f"{tag}_{name}_v"])
        assert np.max(np.abs(v - gold[f"{tag}_{name}_v"]) / scale) <= 1e-12
        assert np.max(np.abs(u - gold[f"{tag}_{name}_u"]) / (1 + np.abs(u))) <= 1e-12
        assert np.array_equal(np.all(v == 0, axis=1), np.all(gold[f"{tag}_{name}_v"] == 0, axis=1))  # same |r| branch
    c.close()


def test_fresh_states_against_the_oracle_including_ragged_batch_sizes(built):
    from armour_b200 import RobustController
    from oracle.pycontroller import OracleController
    c, o = RobustController(MODEL, 0.03), OracleController(eps=0.03)
    rng = np.random.default_rng(3)
    for n in (1, 127, 129, 300):
        q = rng.uniform(-np.pi, np.pi, (n, 7))
        qd, qda, qdd = (rng.uniform(-2, 2, (n, 7)) for _ in range(3))
        qd[:: 9] = 0.0
        for friction in (False, True):
            tau, lo, hi = c.rnea(q, qd, qda, qdd, friction=friction)
            for i in range(0, n, max(1, n // 40)):
                rlo, rhi = o.rnea_interval(q[i], qd[i], qda[i], qdd[i], friction=friction)
                assert np.array_equal(lo[i], rlo) and np.array_equal(hi[i], rhi)
                assert np.max(np.abs(tau[i] - o.rnea(q[i], qd[i], qda[i], qdd[i], friction=friction))) <= 1e-12
            assert np.all(lo <= tau) and np.all(tau <= hi)  # the nominal model is one of the models of the interval model
    c.close()


def test_device_pointer_calls_with_device_trigonometry(built):
    import torch

    from armour_b200 import RobustController
    c = RobustController(MODEL, 0.03)
    rng = np.random.default_rng(4)
    n = 4096
    q = rng.uniform(-np.pi, np.pi, (n, 7))
    qd, qda, qdd = (rng.uniform(-2, 2, (n, 7)) for _ in range(3))
    tau_h, lo_h, hi_h = c.rnea(q, qd, qda, qdd)
    dev = torch.device("cuda", 0)
    t = [torch.tensor(a, dtype=torch.float64, device=dev) for a in (q, qd, qda, qdd)]
    out = [torch.empty((n, 7), dtype=torch.float64, device=dev) for _ in range(3)]
    torch.cuda.synchronize()
    c.rnea_device(n, *(x.data_ptr() for x in t), d_tau=out[0].data_ptr(), d_tau_lo=out[1].data_ptr(), d_tau_hi=out[2].data_ptr())
    c.synchronize()
    tau_d, lo_d, hi_d = (x.cpu().numpy() for x in out)
    for a, b in ((tau_d, tau_h), (lo_d, lo_h), (hi_d, hi_h)):
        assert np.max(np.abs(a - b) / (1 + np.abs(b))) <= 1e-12
    # with the host's sin / cos handed over the device-pointer call is the host-pointer call
    sc = torch.tensor(np.stack([np.sin(-q), np.cos(-q)], axis=-1), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    c.rnea_device(n, *(x.data_ptr() for x in t), d_tau_lo=out[1].data_ptr(), d_tau_hi=out[2].data_ptr(), d_sincos=sc.data_ptr())
    c.synchronize()
    # numpy's sin / cos are not glibc's; the end points then agree to rounding, not to the bit
    assert np.max(np.abs(out[1].cpu().numpy() - lo_h) / (1 + np.abs(lo_h))) <= 1e-12
    # controller update on device pointers
    q_des, qd_des, qdd_des = q + rng.uniform(-0.05, 0.05, (n, 7)), qd + rng.uniform(-0.1, 0.1, (n, 7)), rng.uniform(-2, 2, (n, 7))
    Kr = np.full(7, 10.0)
    u_h, un_h, v_h, st_h = c.update(Kr, 1.0, 1e-2, 1e-10, q, qd, q_des, qd_des, qdd_des)
    td = [torch.tensor(a, dtype=torch.float64, device=dev) for a in (q, qd, q_des, qd_des, qdd_des)]
    u_d = torch.empty((n, 7), dtype=torch.float64, device=dev)
    st_d = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    c.update_device(n, Kr, 1.0, 1e-2, 1e-10, *(x.data_ptr() for x in td), d_u=u_d.data_ptr(), d_status=st_d.data_ptr())
    c.synchronize()
    assert np.max(np.abs(u_d.cpu().numpy() - u_h) / (1 + np.abs(u_h))) <= 1e-10
    assert np.array_equal(st_d.cpu().numpy(), st_h) and not st_h.any()
    c.close()


def test_argument_errors(built):
    from armour_b200 import ArmourError, RobustController
    with pytest.raises(ArmourError):
        RobustController("/no/such/model.txt")
    with pytest.raises(ArmourError):
        RobustController(MODEL, model_uncertainty=1.5)
    c = RobustController(MODEL)
    with pytest.raises(ValueError):
        c.rnea(np.zeros((2, 6)), np.zeros((2, 6)), np.zeros((2, 6)), np.zeros((2, 6)))
    with pytest.raises(ArmourError):
        c.rnea(np.zeros((1, 7)), np.zeros((1, 7)), np.zeros((1, 7)), np.zeros((1, 7)), nominal=False, interval=False)
    c.close()
