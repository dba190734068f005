import sys; sys.path.insert(0,'.')
import numpy as np, time
from armour_b200 import ReachSetEngine, worlds, ArmourError
q0, qd0, qdd0, _, obs = worlds.random_problems(1, 10, seed=4)
for thr in (5e-4, 2e-4, 1e-4, 5e-5, 5e-6):
    try:
        eng = ReachSetEngine(max_problems=1, max_obstacles=10, simplify_threshold=thr, cap_link=256, cap_torque=512, cap_work=4096)
        t0=time.time(); eng.build(q0[0], qd0[0], qdd0[0], obs[0]); eng.synchronize(); dt=time.time()-t0
        ln, un = eng.monomial_counts()
        print(f"thr {thr:g}: ok build {dt*1e3:.1f} ms, max link monos {ln.max()}, max torque monos {un.max()}")
    except ArmourError as e:
        print(f"thr {thr:g}: {e}")
