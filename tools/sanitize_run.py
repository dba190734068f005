"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): two planning problems through the latency
build (K1 + K3a as programmatic dependent), evaluations, the batched solver, and the controller kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from armour_b200 import ReachSetEngine, RobustController, worlds  # noqa: E402

q0, qd0, qdd0, q_des, obs = worlds.config1_problem(os.path.join(ROOT, "tests", "golden", "worlds", "scene_016_006.csv"))
eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
eng.build(q0, qd0, qdd0, obs)            # latency configuration: K3a under the tail of K1
g, j = eng.eval(np.zeros(7))
print("single:", float(np.abs(g).max()))
eng.close()
Q0, QD0, QDD0, QDES, OBS = worlds.random_problems(3, 10, seed=5)
eng = ReachSetEngine(max_problems=3, max_obstacles=10)
eng.build(Q0, QD0, QDD0, OBS)            # throughput configuration
g, j = eng.eval(np.zeros((3, 7)))
k, ok, first, it = eng.solve(QDES, max_iter=6)
print("batch:", float(np.abs(g).max()), ok.tolist(), it.tolist())
eng.close()
c = RobustController(os.path.join(ROOT, "tests", "golden", "robot_models", "kinova_without_gripper.txt"))
rng = np.random.default_rng(0)
a = [rng.uniform(-1, 1, (200, 7)) for _ in range(5)]
u, un, v, st = c.update(np.full(7, 10.0), 1.0, 1e-2, 1e-10, *a)
print("controller:", float(np.abs(u).max()), int(st.sum()))
c.close()
