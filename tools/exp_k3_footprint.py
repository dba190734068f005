"""Footprint experiment for k_constraints: the same launch with smaller table capacities (cap_link / cap_torque are
context options; HP_CAP is a build option of the library selected by ARMOUR_B200_LIB).  Prints one line per case."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from armour_b200 import ReachSetEngine, worlds
tag, cap_link, cap_torque = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nprob, iters = 1024, 8
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
for nobs in (10, 0):
    q0, qd0, qdd0, _, obs = worlds.random_problems(nprob, max(nobs, 1))
    obs = obs[:, :nobs]
    eng = ReachSetEngine(max_problems=nprob, max_obstacles=max(nobs, 1), cap_link=cap_link, cap_torque=cap_torque)
    eng.set_stream(st.cuda_stream)
    t = [torch.tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev) for x in (q0, qd0, qdd0, obs)]
    eng.build_device(nprob, nobs, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr() if nobs else t[0].data_ptr())
    torch.cuda.synchronize()
    bad = int((eng.build_status() != 0).sum())
    eng.nobs = nobs
    ks = torch.tensor(worlds.halton_k(iters * nprob).reshape(iters, nprob, 7), dtype=torch.float64, device=dev)
    g = torch.empty((nprob, eng.m), dtype=torch.float64, device=dev)
    j = torch.empty((nprob, eng.m, 7), dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, gp, jp in (("g+jac", g.data_ptr(), j.data_ptr()), ("none", 0, 0)):
        for rep in range(2):
            e0.record(st)
            for it in range(iters):
                eng.eval_device(nprob, ks[it].data_ptr(), gp, jp)
            e1.record(st); torch.cuda.synchronize()
        print(f"{tag} caps {cap_link}/{cap_torque} nobs={nobs:2d} {name:6s}: {1e3*e0.elapsed_time(e1)/iters:8.1f} us per launch"
              f" | overflowed builds {bad} | checksum g {float(g.sum()):.12e}")
    eng.close()
