"""Generate tests/golden/ref/*.npz from the REFERENCE's own sources (oracle/_ref, built by oracle/Makefile.ref
from /root/reference with the stand-in Eigen / Boost.Interval headers of oracle/ref_shim/).

Run in the build container (needs /root/reference):   python tools/make_golden.py
Each fixture holds, for one planning problem and a subset of time intervals, the reach-set tables the
reference planner has after sections II.A-II.C of main() (KPR/armour_main.cu:96-201) and the rows its
eval_g / eval_jac_g compute on the host at two k (KPR/NLPclass.cu:304-320, 376-391).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from armour_b200 import worlds  # noqa: E402  (input generators only)
from oracle import pyref  # noqa: E402

NF = 7
T_SUBSET = list(range(0, 128, 8)) + [127]
KS = np.array([[0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7],  # KPR/PZ_tests.cu:198
               [-0.25, 0.9, -1.0, 0.3, 1.0, -0.8, 0.05]])


def problems():
    out = {"scene_016_006": worlds.config1_problem(os.path.join(ROOT, "tests", "golden", "worlds", "scene_016_006.csv"))}
    q0, qd0, qdd0, qdes, obs = worlds.random_problems(2, 10, seed=5)
    for p in range(2):
        out[f"random_seed5_{p}"] = (q0[p], qd0[p], qdd0[p], qdes[p], obs[p])
    # the debug initial condition of KPR/debug_script.m:29-31
    out["debug_script"] = (-np.ones(NF), np.ones(NF), 2 * np.ones(NF), np.zeros(NF), np.zeros((0, 12)))
    return out


def pack(ref, q0, qd0, qdd0):
    NJ = ref.NJ
    tb = ref.tables(64, 128)
    sel_l = np.array([t * NJ + l for t in T_SUBSET for l in range(NJ)])
    sel_u = np.array([t * NF + j for t in T_SUBSET for j in range(NF)])
    d = dict(q0=q0, qd0=qd0, qdd0=qdd0, t_subset=np.array(T_SUBSET), ks=KS,
             nl=tb["nl"][sel_l], cl=tb["cl"][sel_l], nu=tb["nu"][sel_u], cu=tb["cu"][sel_u], ru=tb["ru"][sel_u],
             torque_radius=tb["torque_radius"][:, T_SUBSET], link_gens=tb["link_gens"][T_SUBSET])
    d["hl"] = np.concatenate([tb["hl"][i, :tb["nl"][i]] for i in sel_l]).astype(np.uint16)
    d["gl"] = np.concatenate([tb["gl"][i, :tb["nl"][i]] for i in sel_l])
    d["hu"] = np.concatenate([tb["hu"][i, :tb["nu"][i]] for i in sel_u]).astype(np.uint16)
    d["gu"] = np.concatenate([tb["gu"][i, :tb["nu"][i]] for i in sel_u])
    for n, k in enumerate(KS):
        s = ref.slice(k)
        d[f"g_torque_{n}"] = s["g_torque"][sel_u]
        d[f"jac_torque_{n}"] = s["jac_torque"][sel_u]
        d[f"link_c_{n}"] = s["link_c"][T_SUBSET]
        d[f"dlink_c_{n}"] = s["dlink_c"][T_SUBSET]
        d[f"bez_{n}"] = s["bez"]
        d[f"dbez_{n}"] = s["dbez"]
    return d


if __name__ == "__main__":
    if not pyref.available():
        raise SystemExit("oracle/_ref is not built and /root/reference is absent")
    out_dir = os.path.join(ROOT, "tests", "golden", "ref")
    os.makedirs(out_dir, exist_ok=True)
    for name, (q0, qd0, qdd0, _, _) in problems().items():
        ref = pyref.ReferenceProblem(q0, qd0, qdd0)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **pack(ref, q0, qd0, qdd0))
        print(name, os.path.getsize(path), "bytes")
