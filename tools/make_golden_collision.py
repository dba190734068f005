"""Generate tests/golden/refcuda/*.npz from the REFERENCE's complete planner path, CUDA kernels included.

oracle/_ref/libarmour_ref_cuda.so = the reference's own PZsparse.cu, Trajectory.cu, Dynamics.cu,
CollisionChecking.cu and NLPclass.cu compiled by nvcc with the reference's flags (oracle/Makefile.ref `cuda`).
It needs a GPU, so this script runs on the GPU box (the library is built in the container and travels):

    make -C oracle -f Makefile.ref cuda
    gpurun -- 'python tools/make_golden_collision.py --out gpurun_out/refcuda'
    cp gpurun_out/refcuda/*.npz tests/golden/refcuda/ ; cp gpurun_out/refcuda/report.json profiles/...

Each fixture holds, for one planning problem (obstacles included) and a k schedule: all of g (torque rows,
collision rows, Bezier rows), the Jacobian on a subset of intervals (every row class), get_bounds_info's g_l / g_u /
x_l / x_u, eval_f / eval_grad_f and finalize_solution's verdict — every number computed by the reference's own code.
Beside the fixtures it writes report.json: the same quantities from the restated oracle (oracle/liboracle.so) compared
with the reference on the spot, including how many collision rows pick a different half-space (argmax) because the
reference's kernels are built with FMA contraction (nvcc default -fmad=true) and the restatement is not.
"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

_spec = importlib.util.spec_from_file_location("armour_worlds", os.path.join(ROOT, "armour_b200", "worlds.py"))
worlds = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(worlds)  # input generators only; does not load the product library

from oracle import pyrefcuda  # noqa: E402
from oracle.pyoracle import OracleProblem  # noqa: E402

NF, T, NJ = 7, 128, 7
T_SUBSET = list(range(0, 128, 8)) + [127]
WORLD_DIR = os.path.join(ROOT, "tests", "golden", "worlds")


def k_schedule():
    """k = 0, the PZ_tests.cu:198 point, Halton points, one corner of the box."""
    ks = [np.zeros(NF), np.array([0.5, 0.6, 0.7, 0.0, -0.5, -0.6, -0.7])]
    ks += list(worlds.halton_k(3, skip=3))
    ks.append(np.array([1.0, -1.0, 1.0, 1.0, -1.0, 1.0, -1.0]))
    return np.array(ks)


def problems():
    out = {}
    for name in ("scene_016_006", "scene_013_001", "scene_028_003", "scene_040_010"):
        out[name] = worlds.config1_problem(os.path.join(WORLD_DIR, name + ".csv"))
    # two worlds of the bench batch (config 2, seed 20261017: the first two problems of the 1 024)
    q0, qd0, qdd0, qdes, obs = worlds.random_problems(2, 10, seed=20261017)
    for p in range(2):
        out[f"bench_seed20261017_{p}"] = (q0[p], qd0[p], qdd0[p], qdes[p], obs[p])
    # a moving start among 40 obstacles (= MAX_OBSTACLE_NUM, KPR/Parameters.h:26)
    q0, qd0, qdd0, qdes, obs = worlds.random_problems(1, 40, seed=40)
    out["random_seed40_40obs"] = (q0[0], qd0[0], qdd0[0], qdes[0], obs[0])
    # the debug initial condition of KPR/debug_script.m:29-31 among the obstacles of scene_016_006
    o = out["scene_016_006"][4]
    out["debug_script_obs"] = (-np.ones(NF), np.ones(NF), 2 * np.ones(NF), np.zeros(NF), o)
    return out


def rows_subset(nobs):
    """Row indices of g whose Jacobian rows are stored: every row class on the intervals of T_SUBSET."""
    rows = [t * NF + j for t in T_SUBSET for j in range(NF)]
    rows += [NF * T + (l * T + t) * nobs + o for l in range(NJ) for t in T_SUBSET for o in range(nobs)]
    m = NF * T + NJ * T * nobs + 4 * NF
    rows += list(range(m - 4 * NF, m))
    return np.array(rows)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "refcuda"))
    ap.add_argument("--product", action="store_true", help="also compare libarmour_b200.so with the reference")
    a = ap.parse_args()
    if not pyrefcuda.available():
        raise SystemExit("oracle/_ref/libarmour_ref_cuda.so missing or no CUDA device")
    os.makedirs(a.out, exist_ok=True)
    ks = k_schedule()
    report = {}
    for name, (q0, qd0, qdd0, q_des, obs) in problems().items():
        ref = pyrefcuda.ReferencePlanner(q0, qd0, qdd0, q_des, obs)
        orc = OracleProblem().build(q0, qd0, qdd0, obs)
        assert ref.m == orc.m
        nobs = ref.nobs
        rs = rows_subset(nobs)
        xl, xu, gl, gu = ref.bounds()
        d = dict(q0=q0, qd0=qd0, qdd0=qdd0, q_des=q_des, obstacles=obs, ks=ks, t_subset=np.array(T_SUBSET), jac_rows=rs,
                 x_l=xl, x_u=xu, g_l=gl, g_u=gu)
        ogl, ogu = orc.bounds()
        rep = dict(m=int(ref.m), nobs=int(nobs), bounds_max_abs_diff=float(max(np.max(np.abs(gl - ogl)), np.max(np.abs(gu - ogu)))),
                   per_k=[])
        # half-spaces: reference kernels vs restatement
        Ar, dr, der = ref.hyperplanes()
        Ao, do, deo = orc.hyperplanes()
        rep["hyperplanes"] = dict(max_abs_dA=float(np.max(np.abs(Ar - Ao))), max_abs_dd=float(np.max(np.abs(dr - do))),
                                  max_abs_ddelta=float(np.max(np.abs(der - deo))),
                                  frac_A_bit_identical=float(np.mean(Ar == Ao)),
                                  frac_d_bit_identical=float(np.mean(dr == do)),
                                  frac_delta_bit_identical=float(np.mean(der == deo)))
        c0, c1 = NF * T, NF * T + NJ * T * nobs
        for n, k in enumerate(ks):
            g, J = ref.eval_g(k), ref.eval_jac_g(k)
            f, gf = ref.cost(k)
            feas = ref.finalize(k, g, f)
            d[f"g_{n}"] = g
            d[f"jac_{n}"] = J[rs]
            d[f"f_{n}"] = np.array(f)
            d[f"grad_f_{n}"] = gf
            d[f"feasible_{n}"] = np.array(int(feas))
            go, Jo = orc.eval_g(k), orc.eval_jac_g(k)
            ok, first = orc.verdict(go)
            dJ = np.abs(J - Jo)
            # a collision row whose Jacobian differs by more than rounding picked another half-space
            flips = int(np.sum(np.max(dJ[c0:c1], axis=1) > 1e-9))
            rep["per_k"].append(dict(
                k=[float(x) for x in k], feasible_ref=bool(feas), feasible_oracle=bool(ok), oracle_first_violation=int(first),
                max_abs_dg_torque=float(np.max(np.abs(g[:c0] - go[:c0]))), max_abs_dg_collision=float(np.max(np.abs(g[c0:c1] - go[c0:c1]))),
                max_abs_dg_bezier=float(np.max(np.abs(g[c1:] - go[c1:]))), max_abs_dJ_torque=float(np.max(dJ[:c0])),
                max_abs_dJ_collision=float(np.max(dJ[c0:c1])), max_abs_dJ_bezier=float(np.max(dJ[c1:])),
                collision_rows=int(c1 - c0), collision_rows_bit_identical=int(np.sum(g[c0:c1] == go[c0:c1])),
                collision_rows_other_halfspace=flips,
                cost_abs_diff=float(abs(f - orc.cost(q_des, k))), cost_grad_max_abs_diff=float(np.max(np.abs(gf - orc.cost_grad(q_des, k))))))
        if a.product:  # the CUDA product path against the reference, on the spot (the GPU tests do it from the fixtures)
            from armour_b200 import ReachSetEngine
            eng = ReachSetEngine(max_problems=1, max_obstacles=max(nobs, 1))
            eng.build(q0, qd0, qdd0, obs)
            pgl, pgu = eng.get_bounds_info()
            rep["product_bounds_max_abs_diff"] = float(max(np.max(np.abs(gl - pgl[0])), np.max(np.abs(gu - pgu[0]))))
            for n, k in enumerate(ks):
                gp, Jp = eng.eval(k)
                okp, _ = eng.finalize_solution(gp[0])
                gr, Jr = d[f"g_{n}"], ref.eval_jac_g(k)
                dJ = np.abs(Jp[0] - Jr)
                rep["per_k"][n].update(product_max_abs_dg=float(np.max(np.abs(gp[0] - gr))), product_max_abs_dJ=float(np.max(dJ)),
                                       product_rows_other_halfspace=int(np.sum(np.max(dJ[c0:c1], axis=1) > 1e-9)),
                                       product_feasible=bool(okp))
            eng.close()
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **d)
        rep["fixture_bytes"] = os.path.getsize(path)
        report[name] = rep
        worst = max(r["max_abs_dg_collision"] for r in rep["per_k"])
        print(f"{name}: m={ref.m} nobs={nobs} bytes={rep['fixture_bytes']} max|dg_coll|={worst:.3e} "
              f"flips={sum(r['collision_rows_other_halfspace'] for r in rep['per_k'])} "
              f"verdicts={[r['feasible_ref'] for r in rep['per_k']]}", flush=True)
        del ref
    with open(os.path.join(a.out, "report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
