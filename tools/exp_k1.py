"""Tuning experiment: build latency (1 problem) and throughput (batch) of the reach-set kernel for the library
selected by ARMOUR_B200_LIB.  Prints one line; not part of the product or the bench."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nprob", type=int, default=256)
ap.add_argument("--nobs", type=int, default=10)
ap.add_argument("--tag", default="")
a = ap.parse_args()
from armour_b200 import ReachSetEngine, worlds  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)


def run(nprob, reps):
    q0, qd0, qdd0, _, obs = worlds.random_problems(nprob, a.nobs)
    eng = ReachSetEngine(max_problems=nprob, max_obstacles=a.nobs)
    eng.set_stream(st.cuda_stream)
    t = [torch.tensor(x, dtype=torch.float64, device=dev) for x in (q0, qd0, qdd0, obs)]
    ms = []
    for r in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.build_device(nprob, a.nobs, *(x.data_ptr() for x in t))
        e1.record(st)
        torch.cuda.synchronize()
        if r >= 2:
            ms.append(e0.elapsed_time(e1))
    assert not eng.build_status().any()
    g, _ = eng.eval(np.zeros((nprob, 7)), True, False)
    tr = eng.torque_radius()
    eng.close()
    return float(np.median(ms)), float(g.sum()), float(tr.sum())


lat, cs1, tr1 = run(1, 10)
thr, csn, trn = run(a.nprob, 3)
print(f"{a.tag or os.environ.get('ARMOUR_B200_LIB', 'default')}: single build {lat:.3f} ms | batch {a.nprob}: {thr:.1f} ms "
      f"= {1e3 * thr / a.nprob:.1f} us/problem | checksums g {cs1:.12e} {csn:.12e} tr {tr1:.12e}")
