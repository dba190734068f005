"""Developer tool: which work capacity does a SIMPLIFY_THRESHOLD need?  Prints the build status per capacity."""
import sys

sys.path.insert(0, ".")
from armour_b200 import ArmourError, ReachSetEngine, worlds  # noqa: E402

thr = float(sys.argv[1]) if len(sys.argv) > 1 else 5e-6
q0, qd0, qdd0, _, obs = worlds.random_problems(1, 100, seed=4)
for cap_work in (8192, 16384, 32768, 65536):
    for cl, cu in ((128, 256),):
        try:
            eng = ReachSetEngine(max_problems=1, max_obstacles=100, simplify_threshold=thr, cap_link=cl, cap_torque=cu,
                                 cap_work=cap_work)
            eng.build(q0[0], qd0[0], qdd0[0], obs[0])
            ln, un = eng.monomial_counts()
            print(thr, cap_work, "ok", ln.max(), un.max(), flush=True)
            eng.close()
        except ArmourError as exc:
            print(thr, cap_work, "FAILED:", exc, flush=True)
