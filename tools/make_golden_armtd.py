"""Freeze outputs of the REFERENCE's ARMTD comparison planner (oracle/_ref/libarmour_ref_armtd.so: the KPA sources compiled by
nvcc, oracle/Makefile.ref target `armtd`) as tests/golden/armtd/reference.npz.  Needs a GPU (the reference's own collision
kernels run):   gpurun -- 'python tools/make_golden_armtd.py'   then copy gpurun_out/armtd_reference.npz to tests/golden/armtd/.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from armour_b200 import worlds  # noqa: E402
from oracle.pyrefarmtd import ReferenceArmtd  # noqa: E402

WORLDS = os.path.join(ROOT, "tests", "golden", "worlds")
CASES = [("scene_016_006.csv", 1), ("scene_013_001.csv", 2), ("scene_022_001.csv", 3), ("scene_016_006.csv", 4)]
BLOCKED = 3  # this case gets an obstacle moved onto link 5 at mid-horizon, so that verdicts with a violated row are pinned too


def k_schedule(seed):
    rng = np.random.default_rng(seed)
    return np.vstack([np.zeros(7), np.array([0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7]), rng.uniform(-1, 1, (3, 7))])


def main():
    out = {}
    for ci, (name, seed) in enumerate(CASES):
        q0, qd0, q_des, jrs, k_range, obs = worlds.armtd_problem(os.path.join(WORLDS, name), seed)
        ref = ReferenceArmtd(q0, qd0, q_des, jrs, k_range, obs)
        if ci == BLOCKED:
            ref.eval_g(np.zeros(7))
            obs = obs.copy()
            obs.reshape(-1, 12)[0, :3] = ref.link_sliced_center()[50, 4]
            ref = ReferenceArmtd(q0, qd0, q_des, jrs, k_range, obs)
        tag = f"c{ci}_"
        out.update({tag + "q0": q0, tag + "qd0": qd0, tag + "q_des": q_des, tag + "jrs": jrs, tag + "k_range": k_range,
                    tag + "obs": obs, tag + "world": np.array(name)})
        gl, gu = ref.bounds()
        out.update({tag + "gl": gl, tag + "gu": gu, tag + "link_gens": ref.link_gens()})
        ks = k_schedule(seed)
        out[tag + "k"] = ks
        G, J, F, DF, V, LC = [], [], [], [], [], []
        for k in ks:
            g = ref.eval_g(k)
            G.append(g)
            LC.append(ref.link_sliced_center())
            J.append(ref.eval_jac_g(k))
            f, df = ref.cost(k)
            F.append(f)
            DF.append(df)
            V.append(ref.finalize(k, g))
        out.update({tag + "g": np.array(G), tag + "J": np.array(J), tag + "f": np.array(F), tag + "df": np.array(DF),
                    tag + "feasible": np.array(V), tag + "link_sliced_center": np.array(LC)})
        tabs = ref.link_tables()
        nmax = max(len(t[1]) for t in tabs)
        n = np.array([len(t[1]) for t in tabs], dtype=np.int32)
        keys = np.zeros((len(tabs), nmax), dtype=np.uint64)
        coeff = np.zeros((len(tabs), nmax, 3))
        cen = np.array([t[0] for t in tabs])
        for i, t in enumerate(tabs):
            keys[i, :n[i]] = t[1]
            coeff[i, :n[i]] = t[2]
        out.update({tag + "tab_n": n, tag + "tab_key": keys, tag + "tab_coeff": coeff, tag + "tab_center": cen})
        print(name, "m =", ref.m, "monomials max", nmax, "feasible", V, "max|g|", float(np.abs(np.array(G)).max()))
    dst = os.path.join(ROOT, "gpurun_out")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "armtd_reference.npz"), **out)
    print("wrote", os.path.join(dst, "armtd_reference.npz"))


if __name__ == "__main__":
    main()
