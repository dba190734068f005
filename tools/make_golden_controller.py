"""Freeze outputs of the REFERENCE's own robust-controller sources (oracle/_ref/libarmour_ref_controller.so, built by
oracle/Makefile.ref target `mex` from MEX/*.cpp) as tests/golden/controller/reference.npz, so that the oracle and the
product can be checked against the reference where /root/reference does not exist.  Run here (CPU only):
    python tools/make_golden_controller.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pycontroller import ReferenceController  # noqa: E402


def states(rng, n):
    """sampled states: positions over the joint ranges, moderate rates; every fifth state almost on its reference (small r),
    some with zero rates"""
    q = rng.uniform(-np.pi, np.pi, (n, 7))
    qd = rng.uniform(-1.5, 1.5, (n, 7))
    q_des = q + rng.uniform(-0.05, 0.05, (n, 7))
    qd_des = qd + rng.uniform(-0.1, 0.1, (n, 7))
    qdd_des = rng.uniform(-2.0, 2.0, (n, 7))
    q_des[::5] = q[::5] + rng.uniform(-1e-9, 1e-9, (len(q[::5]), 7))
    qd_des[::5] = qd[::5]
    qd[3::7] = 0.0
    q_des[4::11] += 2 * np.pi  # wrap-around of the position error
    return q, qd, q_des, qd_des, qdd_des


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    n = 64
    q, qd, q_des, qd_des, qdd_des = states(rng, n)
    qda = rng.uniform(-1.5, 1.5, (n, 7))
    qdd = rng.uniform(-3.0, 3.0, (n, 7))
    out.update(q=q, qd=qd, qda=qda, qdd=qdd, q_des=q_des, qd_des=qd_des, qdd_des=qdd_des)
    for eps in (0.03, 0.0, 0.1):
        ref = ReferenceController(eps=eps)
        tag = f"eps{eps}"
        out[tag + "_model"] = ref.interval_model()
        tau = np.empty((n, 7)); lo = np.empty((n, 7)); hi = np.empty((n, 7))
        lo_ng = np.empty((n, 7)); hi_ng = np.empty((n, 7))
        for i in range(n):
            tau[i] = ref.rnea(q[i], qd[i], qda[i], qdd[i])
            lo[i], hi[i] = ref.rnea_interval(q[i], qd[i], qda[i], qdd[i])
            lo_ng[i], hi_ng[i] = ref.rnea_interval(q[i], qd[i], qda[i], qdd[i], gravity=False)
        out.update({tag + "_tau": tau, tag + "_lo": lo, tag + "_hi": hi, tag + "_lo_nograv": lo_ng, tag + "_hi_nograv": hi_ng})
        Kr = np.array([10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0])
        out["Kr"] = Kr
        for name, (alpha, V_max, thr) in {"a": (1.0, 1e-2, 1e-10), "b": (20.0, 1e-5, 1e-7)}.items():
            u = np.empty((n, 7)); un = np.empty((n, 7)); v = np.empty((n, 7)); st = np.empty(n, dtype=np.int32)
            for i in range(n):
                u[i], un[i], v[i], st[i] = ref.update(Kr, alpha, V_max, thr, q[i], qd[i], q_des[i], qd_des[i], qdd_des[i])
            out.update({f"{tag}_{name}_gains": np.array([alpha, V_max, thr]), f"{tag}_{name}_u": u, f"{tag}_{name}_un": un,
                        f"{tag}_{name}_v": v, f"{tag}_{name}_status": st})
    dst = os.path.join(ROOT, "tests", "golden", "controller")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "reference.npz"), **out)
    print("wrote", os.path.join(dst, "reference.npz"), {k: v.shape for k, v in out.items() if k.startswith("eps0.03")})


if __name__ == "__main__":
    main()
