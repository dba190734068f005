"""Tuning experiment: time of one k_constraints launch for a resident batch (device pointers)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from armour_b200 import ReachSetEngine, worlds  # noqa: E402

nprob, nobs, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 10, 16
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
q0, qd0, qdd0, _, obs = worlds.random_problems(nprob, nobs)
eng = ReachSetEngine(max_problems=nprob, max_obstacles=nobs)
eng.set_stream(st.cuda_stream)
t = [torch.tensor(x, dtype=torch.float64, device=dev) for x in (q0, qd0, qdd0, obs)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
eng.build_device(nprob, nobs, *(x.data_ptr() for x in t))
e1.record(st)
torch.cuda.synchronize()
build_ms = e0.elapsed_time(e1)
ks = torch.tensor(worlds.halton_k(iters * nprob).reshape(iters, nprob, 7), dtype=torch.float64, device=dev)
g = torch.empty((nprob, eng.m), dtype=torch.float64, device=dev)
j = torch.empty((nprob, eng.m, 7), dtype=torch.float64, device=dev)
for rep in range(3):
    e0.record(st)
    for it in range(iters):
        eng.eval_device(nprob, ks[it].data_ptr(), g.data_ptr(), j.data_ptr())
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
print(f"{sys.argv[2] if len(sys.argv) > 2 else ''}: build {build_ms:.1f} ms ({1e3*build_ms/nprob:.1f} us/problem) | k_constraints {ms*1e3:.1f} us per launch of {nprob} "
      f"| checksum g {float(g.sum()):.12e} J {float(j.sum()):.12e}")
