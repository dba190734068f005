import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from armour_b200 import ReachSetEngine, worlds
nprob, nobs, iters = 1024, 10, 16
dev = torch.device("cuda", 0); st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
q0, qd0, qdd0, _, obs = worlds.random_problems(nprob, nobs)
eng = ReachSetEngine(max_problems=nprob, max_obstacles=nobs); eng.set_stream(st.cuda_stream)
t = [torch.tensor(x, dtype=torch.float64, device=dev) for x in (q0, qd0, qdd0, obs)]
eng.build_device(nprob, nobs, *(x.data_ptr() for x in t)); torch.cuda.synchronize()
ks = torch.tensor(worlds.halton_k(iters * nprob).reshape(iters, nprob, 7), dtype=torch.float64, device=dev)
g = torch.empty((nprob, eng.m), dtype=torch.float64, device=dev)
j = torch.empty((nprob, eng.m, 7), dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, gp, jp in (("g+J", g.data_ptr(), j.data_ptr()), ("g only", g.data_ptr(), 0), ("J only", 0, j.data_ptr())):
    for rep in range(3):
        e0.record(st)
        for it in range(iters):
            eng.eval_device(nprob, ks[it].data_ptr(), gp, jp)
        e1.record(st); torch.cuda.synchronize()
    print(name, f"{e0.elapsed_time(e1)/iters*1e3:.1f} us per launch")
