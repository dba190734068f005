"""Multi-rank host logic on CPU: two processes over gloo shard a batch of planning problems, 'solve' their
own block and gather per-problem results into global order (what bench.py / a batched planner do over NCCL
after the timed region).  No GPU: the per-problem work is stood in for by the oracle on a tiny problem."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from armour_b200 import sharding, worlds
    from oracle.pyoracle import OracleProblem
    lo, hi = sharding.shard_bounds(n_total, world, rank)
    q0, qd0, qdd0, q_des, obs = worlds.random_problems(n_total, 2, seed=99)  # same generator on every rank
    ks = worlds.halton_k(n_total)
    feas, first, ident = [], [], []
    for p in range(lo, hi):
        ref = OracleProblem(num_time_steps=8).build(q0[p], qd0[p], qdd0[p], obs[p], nthreads=1)
        ok, row = ref.verdict(ref.eval_g(ks[p]))
        feas.append(int(ok))
        first.append(row)
        ident.append(p)
    local = {"feasible": torch.tensor(feas, dtype=torch.int32), "first": torch.tensor(first, dtype=torch.int32),
             "id": torch.tensor(ident, dtype=torch.int64), "k": torch.tensor(ks[lo:hi])}
    glob = sharding.gather_results(local, n_total)
    tmax = sharding.reduce_max(10.0 + rank)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), tmax=tmax, **{k: v.numpy() for k, v in glob.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from armour_b200.sharding import shard_bounds
    for n in (0, 1, 5, 8, 1024, 65536 + 3):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gather_results_in_global_order(built, tmp_path):
    n_total, world = 5, 2  # uneven shards: 3 + 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    assert got["tmax"] == 11.0
    assert np.array_equal(got["id"], np.arange(n_total))
    # the gathered verdicts equal a single-process run
    import sys
    sys.path.insert(0, ROOT)
    from armour_b200 import worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, q_des, obs = worlds.random_problems(n_total, 2, seed=99)
    ks = worlds.halton_k(n_total)
    assert np.allclose(got["k"], ks)
    for p in range(n_total):
        ref = OracleProblem(num_time_steps=8).build(q0[p], qd0[p], qdd0[p], obs[p], nthreads=1)
        ok, row = ref.verdict(ref.eval_g(ks[p]))
        assert got["feasible"][p] == int(ok) and got["first"][p] == row
