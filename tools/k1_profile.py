"""Developer tool: per-operation-site cycle profile of the reach-set kernel for ONE planning problem.
Needs a library built with -DK1_PROFILE (exp/lib_prof.so); prints, for the slowest interval, the cycles per
operation site (source line of k1_reachsets.cuh) and the distribution of unit times over the 128 intervals."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ARMOUR_B200_LIB", os.path.join(ROOT, "exp", "lib_prof.so"))
import numpy as np  # noqa: E402

from armour_b200 import ReachSetEngine, worlds  # noqa: E402

SITES = 512
q0, qd0, qdd0, _, obs = worlds.config1_problem(os.path.join(ROOT, "tests", "golden", "worlds", "scene_016_006.csv"))
eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
lib = eng.lib
lib.armour_debug_k1_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
eng.build(q0, qd0, qdd0, obs)
eng.synchronize()
lib.armour_debug_k1_profile(None, 1)
REPS = 5
for _ in range(REPS):
    eng.build(q0, qd0, qdd0, obs)
eng.synchronize()
buf = np.zeros((128, SITES, 2), dtype=np.int64)
lib.armour_debug_k1_profile(buf.ctypes.data_as(ctypes.c_void_p), 0)
cyc = buf[:, :, 0] / REPS
line = buf[:, :, 1] & 0xffffffff
nmax = buf[:, :, 1] >> 32
unit = cyc[:, SITES - 1]
print("unit cycles: min %.0f  median %.0f  max %.0f (t=%d)  sum/128 %.0f" % (unit.min(), np.median(unit), unit.max(), unit.argmax(), unit.mean()))
print("per-interval:", " ".join("%d" % (u / 1000) for u in unit), "(kcycles)")
t = int(unit.argmax())
ops = cyc[t, :SITES - 1]
tot_ops = ops.sum()
print(f"slowest interval t={t}: unit {unit[t]:.0f} cycles, operation sites {tot_ops:.0f} ({100*tot_ops/unit[t]:.1f}%)")
by_line, n_line = {}, {}
for s in range(SITES - 1):
    if ops[s] > 0:
        by_line[int(line[t, s])] = by_line.get(int(line[t, s]), 0) + ops[s]
        n_line[int(line[t, s])] = max(n_line.get(int(line[t, s]), 0), int(nmax[t, s]))
src = open(os.path.join(ROOT, "armour_b200", "csrc", "k1_reachsets.cuh")).read().split("\n")
for ln, c in sorted(by_line.items(), key=lambda kv: -kv[1]):
    print(f"  {100*c/unit[t]:5.1f}%  {c:9.0f}  nmax {n_line[ln]:4d}  L{ln}: {src[ln-1].strip()[:100]}")
