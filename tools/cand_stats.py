"""Developer tool: candidate half-space statistics of a built batch (how many of the 72 half-spaces per row survive)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from armour_b200 import ReachSetEngine, worlds  # noqa: E402

nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 64
q0, qd0, qdd0, _, obs = worlds.random_problems(nprob, 10, seed=20261017)
eng = ReachSetEngine(max_problems=nprob, max_obstacles=10)
eng.build(q0, qd0, qdd0, obs)
c = eng.candidate_counts().astype(np.int64)
print(f"candidates per row: mean {c.mean():.3f}  median {np.median(c):.0f}  p99 {np.percentile(c, 99):.0f}  max {c.max()}  "
      f"overflow rows {(c == 255).sum()}  bytes per problem {c[c < 255].sum() * 32 / nprob:.0f}")
print("histogram:", np.bincount(c.ravel(), minlength=12)[:34].tolist())
