#!/usr/bin/env python
"""BASELINE config 5: batched replanning sweep over 65,536 random worlds (config-2 generator, seed = world index
block), sharded over the GPUs of one box with armour_b200.sharding.shard_bounds; no collective on the data path.

Each rank walks its shard in batches of --batch worlds through ONE context: reach-set build (M1) + --iters
eval_g + eval_jac_g pairs per world (M2) + the device verdict of the last iterate.  Timed on the device with CUDA
events on the launching stream, max over ranks; rank 0 prints one JSON line.  Launch:
    python tools/sweep65k.py                                  (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep65k.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from armour_b200 import ReachSetEngine, sharding, worlds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--worlds", type=int, default=65536)
ap.add_argument("--batch", type=int, default=2048)
ap.add_argument("--nobs", type=int, default=10)
ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)  # the NCCL banner goes to stderr
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)

lo, hi = sharding.shard_bounds(a.worlds, world, rank)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
eng = ReachSetEngine(max_problems=a.batch, max_obstacles=a.nobs, device=local)
eng.set_stream(stream.cuda_stream)
m = eng.lib.armour_num_constraints(eng._h, a.nobs)
d_g = torch.empty((a.batch, m), dtype=torch.float64, device=dev)
d_j = torch.empty((a.batch, m, 7), dtype=torch.float64, device=dev)
d_ok = torch.empty(a.batch, dtype=torch.int32, device=dev)
d_first = torch.empty(a.batch, dtype=torch.int32, device=dev)
ks = torch.tensor(worlds.halton_k(a.iters * a.batch).reshape(a.iters, a.batch, 7), dtype=torch.float64, device=dev)

t_build = t_eval = 0.0
feasible = failed = 0
for b0 in range(lo, hi, a.batch):
    n = min(a.batch, hi - b0)
    # inputs are generated on the host outside the timed regions (seed = first world of the batch)
    q0, qd0, qdd0, _, obs = worlds.random_problems(n, a.nobs, seed=1000003 + b0)
    t = [torch.tensor(x, dtype=torch.float64, device=dev) for x in (q0, qd0, qdd0, obs)]
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    eng.build_device(n, a.nobs, *(x.data_ptr() for x in t))
    e1.record(stream)
    for it in range(a.iters):
        eng.eval_device(n, ks[it].data_ptr(), d_g.data_ptr(), d_j.data_ptr())
    eng.verdict_device(n, d_g.data_ptr(), d_ok.data_ptr(), d_first.data_ptr())
    e2.record(stream)
    torch.cuda.synchronize()
    t_build += e0.elapsed_time(e1)
    t_eval += e1.elapsed_time(e2)
    failed += int((eng.build_status()[:n] != 0).sum())
    feasible += int(d_ok[:n].sum().item())

tb = sharding.reduce_max(t_build, dev)
te = sharding.reduce_max(t_eval, dev)
counts = torch.tensor([feasible, failed, hi - lo], dtype=torch.float64, device=dev)
if dist is not None:
    dist.all_reduce(counts)
if rank == 0:
    print(json.dumps({
        "config": "BASELINE config 5: replanning sweep", "worlds": a.worlds, "n_gpus": world, "batch_per_context": a.batch,
        "obstacles": a.nobs, "k_iterates_per_world": a.iters,
        "build_s": tb * 1e-3, "eval_s": te * 1e-3,
        "m1_problems_per_s": a.worlds / ((tb + te / a.iters) * 1e-3),
        "m2_evals_per_s": a.worlds * a.iters / (te * 1e-3),
        "build_us_per_world_per_gpu": 1e3 * tb / ((a.worlds + world - 1) // world),
        "feasible_last_iterate": int(counts[0].item()), "capacity_failures": int(counts[1].item()),
        "worlds_done": int(counts[2].item()), "timing": "CUDA events on the launching stream, max over ranks"}), flush=True)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
