"""Developer tool: throughput of the batched device solver (plans per second) on the bench batch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from armour_b200 import ReachSetEngine, worlds  # noqa: E402

nprob = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
q0, qd0, qdd0, q_des, obs = worlds.random_problems(nprob, 10, seed=20261017)
eng = ReachSetEngine(max_problems=nprob, max_obstacles=10)
eng.build(q0, qd0, qdd0, obs)
eng.synchronize()
k, ok, first, iters = eng.solve(q_des)  # warm-up (allocations)
l0 = eng.kernel_launches
t0 = time.perf_counter()
k, ok, first, iters = eng.solve(q_des)
dt = time.perf_counter() - t0
print("iterations histogram:", np.bincount(iters, minlength=61).tolist())
print(f"{nprob} problems: solve {dt*1e3:.1f} ms = {nprob/dt:.0f} plans/s; feasible {int(ok.sum())}; iterations mean {iters.mean():.1f} "
      f"max {iters.max()}; constraint evaluations {int((2 * iters + 2).sum())} -> {(2*iters+2).sum()/dt:.0f} evals/s; "
      f"kernel launches {eng.kernel_launches - l0}")
