"""Small driver for ncu captures: builds `nprob` problems and evaluates them `reps` times (no timing here)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nprob", type=int, default=1)
ap.add_argument("--nobs", type=int, default=10)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--world", default="scene_016_006.csv")
a = ap.parse_args()
from armour_b200 import ReachSetEngine, worlds  # noqa: E402

if a.nprob == 1:
    q0, qd0, qdd0, _, obs = worlds.config1_problem(os.path.join(ROOT, "tests", "golden", "worlds", a.world))
    obs = obs[:a.nobs]
else:
    q0, qd0, qdd0, _, obs = worlds.random_problems(a.nprob, a.nobs)
eng = ReachSetEngine(max_problems=a.nprob, max_obstacles=a.nobs)
ks = worlds.halton_k(a.nprob * a.reps).reshape(a.reps, a.nprob, 7)
for r in range(a.reps):
    eng.build(q0, qd0, qdd0, obs)
    g, j = eng.eval(ks[r])
print("ok", eng.m, float(np.abs(g).max()), eng.kernel_launches)
