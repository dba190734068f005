"""Developer tool: task timeline of the reach-set kernel (MG latency configuration) for ONE planning problem.
Needs a library built with -DK1_PROFILE (exp/lib_prof.so).  Prints the per-interval unit times and, for the slowest
interval, every task with its claim time, end time, cycles spent waiting for inputs and the group that ran it."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ARMOUR_B200_LIB", os.path.join(ROOT, "exp", "lib_prof.so"))
import numpy as np  # noqa: E402

from armour_b200 import ReachSetEngine, worlds  # noqa: E402

KINDS = ["W", "WA", "WD", "T4", "LA", "T10", "TF", "TN", "FKC", "FKL", "FB", "NB", "U", "EPI"]
q0, qd0, qdd0, _, obs = worlds.config1_problem(os.path.join(ROOT, "tests", "golden", "worlds", "scene_016_006.csv"))
eng = ReachSetEngine(max_problems=1, max_obstacles=obs.shape[0])
lib = eng.lib
lib.armour_debug_k1_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
for _ in range(3):
    eng.build(q0, qd0, qdd0, obs)
eng.synchronize()
buf = np.zeros((128, 256, 4), dtype=np.int64)
lib.armour_debug_k1_profile(buf.ctypes.data_as(ctypes.c_void_p), 0)
unit = buf[:, 255, 0]
print("unit cycles: min %d median %d max %d (t=%d)" % (unit.min(), np.median(unit), unit.max(), unit.argmax()))
t = int(unit.argmax())
rows = buf[t]
busy = {}
print(f"interval t={t}: task  group  claim  end  run  wait")
for k in range(255):
    c0, c1, w, code = rows[k]
    if c1 == 0:
        continue
    g, kind, i = code >> 16, (code >> 8) & 255, code & 255
    busy[g] = busy.get(g, 0) + (c1 - c0 - w)
    print(f"  {k:3d} {KINDS[kind]:4s}({i}) g{g}  {c0:8d} {c1:8d}  run {c1-c0-w:7d}  wait {w:7d}")
print("busy cycles per group:", busy, " unit:", unit[t])
bykind = {}
for k in range(255):
    c0, c1, w, code = rows[k]
    if c1:
        kind = KINDS[(code >> 8) & 255]
        bykind[kind] = bykind.get(kind, 0) + (c1 - c0 - w)
print("run cycles per task kind:", bykind)
