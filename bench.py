#!/usr/bin/env python
"""bench.py — throughput of the ARMOUR reach-set / constraint-evaluation hot path on B200.

Workload (BASELINE.json configs[1], SURVEY.md 8d "config 2"): per GPU 1,024 random Kinova Gen3 worlds x 10
obstacles (seed 20261017 + rank), 128 time intervals; a STEP is one batch of Ipopt-style queries: for each
of `--iters` (16) k-iterates, one eval_g + one eval_jac_g for every problem of the batch (metric M2,
"planning iters/s": one unit = one eval_g + eval_jac_g pair on built reach sets).  The reach-set build
(metric M1) is timed separately and reported in the "m1" object, next to the single-problem latency
BASELINE.json targets (< 1 ms for one Kinova planning iteration's build + eval).

  value   : M2 units/s, reach sets, k and outputs resident in HBM (device pointers, CUDA events)
  e2e     : M2 units/s through the C-ABI host-buffer call (armour_batch_eval: pinned k -> device,
            kernel, g and dense Jacobian -> pinned host), copies inside the timed region
  roofline: k_constraints, the only kernel of the timed region; algorithmic bytes per launch = B_eval of
            SURVEY.md 8d from the stored monomial counts of THIS batch
  cpu_baseline / --impl reference: the CPU oracle (oracle/, a C++ restatement of the reference: the reference
            itself needs Eigen/Boost/Ipopt and cannot be compiled here) on the box's host cores.

Multi-GPU: problems are independent -> every rank owns its own 1,024 worlds, no collective on the data
path ("weak" scaling); NCCL only reduces the timing (max over ranks) and gathers the verdict counts.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")  # oracle legs parallelise over problems, one thread each

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NF, T = 7, 128
METRIC = "planning iters/s (eval_g + eval_jac_g pairs on built reach sets; Kinova 7-DOF, 128 steps, 10 obstacles)"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nprob", type=int, default=1024, help="worlds per GPU")
    ap.add_argument("--nobs", type=int, default=10)
    ap.add_argument("--iters", type=int, default=16, help="k-iterates per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline eval loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-m1", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, windows):
        """`windows`: (t0, t1) wall-clock pairs of the timed regions, the device-timed one first.  The device
        region of the default run lasts ~50 ms, so when it holds fewer than 3 samples the other timed regions
        of the same run (host-buffer path, M1), equally under load, are added; `window` says which were used."""
        out = self._summary(windows[:1])
        out["window"] = "device-timed region"
        if out["samples"] < 3 and len(windows) > 1:
            out = self._summary(windows)
            out["window"] = "all timed regions of the run (device, host-buffer, M1)"
        return out

    def _summary(self, windows):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if not any(t0 <= ts <= t1 + 0.05 for t0, t1 in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU oracle legs (the checker timed as the CPU baseline; never on the product path)
def oracle_run(nthreads, nobs, iters, steps, warmup, seed, seconds=None):
    """`nthreads` problems, one host thread each (ctypes releases the GIL; the oracle is thread-safe).
    Returns build seconds (all problems, in parallel) and the per-step seconds of `iters` eval pairs per
    problem.  With `seconds` set, runs steps until the budget is spent (at least one)."""
    from concurrent.futures import ThreadPoolExecutor

    from armour_b200 import worlds  # input generator only (pure numpy); no device work
    from oracle.pyoracle import OracleProblem, lib
    lib()
    q0, qd0, qdd0, _, obs = worlds.random_problems(nthreads, nobs, seed=seed)
    ks = worlds.halton_k(iters * max(1, nthreads)).reshape(iters, -1, NF)
    probs = [OracleProblem(max_obstacles=max(40, nobs)) for _ in range(nthreads)]
    pool = ThreadPoolExecutor(nthreads)
    t0 = time.perf_counter()
    list(pool.map(lambda i: probs[i].build(q0[i], qd0[i], qdd0[i], obs[i], nthreads=1), range(nthreads)))
    build_s = time.perf_counter() - t0

    def one(i):
        for it in range(iters):
            probs[i].eval_g(ks[it, i])
            probs[i].eval_jac_g(ks[it, i])

    times = []
    n = 0
    tstart = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        list(pool.map(one, range(nthreads)))
        dt = time.perf_counter() - t0
        n += 1
        if n > warmup:
            times.append(dt)
        if seconds is None:
            if len(times) >= steps:
                break
        elif len(times) >= 1 and time.perf_counter() - tstart > seconds:
            break
    pool.shutdown()
    return build_s, times


def reference_build_timing(nobs, seed):
    """Reach-set build of ONE world by the REFERENCE's own sources (oracle/_ref: KPR PZsparse / Trajectory / Dynamics
    compiled against stand-in Eigen / Boost headers, OpenMP over the 128 intervals like the reference).  None if the
    library did not travel to this box."""
    try:
        from oracle import pyref
        if not os.path.exists(pyref.LIB_PATH):
            return None
        from armour_b200 import worlds
        q0, qd0, qdd0, _, _ = worlds.random_problems(2, max(nobs, 1), seed=seed)
        cores = os.cpu_count() or 1
        best = None
        for p in range(2):
            t0 = time.perf_counter()
            pyref.ReferenceProblem(q0[p], qd0[p], qdd0[p], nthreads=cores)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return {"build_s_per_world": best, "threads": cores, "worlds_per_s": 1.0 / best,
                "kind": "reference sources (KPR/PZsparse.cu, Trajectory.cu, Dynamics.cu) + stand-in Eigen/Boost headers, "
                        "OpenMP over intervals; sections II.A-II.C of main() without the CUDA hyper-plane stage"}
    except Exception as exc:  # the baseline must never take the bench down
        return {"error": repr(exc)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    build_s, times = oracle_run(cores, args.nobs, args.iters, args.steps, args.warmup, seed=20261017)
    total = sum(times)
    units = cores * args.iters * len(times)
    value = units / total
    sample = (f"{cores} worlds (one per host thread, config-2 generator seed 20261017) x {args.iters} k-iterates per "
              f"step; oracle = C++ restatement of the reference (g++ -O2, no Eigen/Boost/Ipopt in the image)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Kinova Gen3 batched random worlds x {args.nobs} obstacles, one eval_g + eval_jac_g "
                               f"per world per k-iterate, {args.iters} k-iterates per step (CPU sample: {cores} worlds)",
                   "time_intervals": T, "obstacles": args.nobs, "k_iterates_per_step": args.iters},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "m1": {"build_s_per_world_1thread": build_s, "builds_per_s_all_cores": cores / build_s,
               "reference_build": reference_build_timing(args.nobs, 20261017)},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    from armour_b200 import ReachSetEngine, worlds
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator is created (NCCL_DEBUG=VERSION/WARN boxes):
        # stdout must carry exactly one JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device=torch.device("cuda", local))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local)
    nprob, nobs, iters = args.nprob, args.nobs, args.iters

    # ---- inputs (synthetic, config-2 generator), reach sets built once per problem -------------------
    q0, qd0, qdd0, q_des, obs = worlds.random_problems(nprob, nobs, seed=20261017 + rank)
    eng = ReachSetEngine(max_problems=nprob, max_obstacles=nobs, device=local)
    # the library launches on THIS stream and the CUDA events below are recorded on it (a NULL handle would
    # make the context fall back to its own stream, which torch events do not see)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    tq0, tqd0, tqdd0, tobs = (torch.tensor(a, dtype=torch.float64, device=dev) for a in (q0, qd0, qdd0, obs))
    launches0 = eng.kernel_launches
    eng.build_device(nprob, nobs, tq0.data_ptr(), tqd0.data_ptr(), tqdd0.data_ptr(), tobs.data_ptr())
    torch.cuda.synchronize()
    status = eng.build_status()
    if status.any():
        raise SystemExit(f"reach-set build overflowed a monomial table for {int((status != 0).sum())} problems")
    m = eng.m
    ks_host = worlds.halton_k(iters * nprob).reshape(iters, nprob, NF)
    d_k = torch.tensor(ks_host, dtype=torch.float64, device=dev)
    d_g = torch.empty((nprob, m), dtype=torch.float64, device=dev)
    d_j = torch.empty((nprob, m, NF), dtype=torch.float64, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        for it in range(iters):
            eng.eval_device(nprob, d_k[it].data_ptr(), d_g.data_ptr(), d_j.data_ptr())

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        timed.launches = eng.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        timed.launches = eng.kernel_launches - timed.launches
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    clocks = ClockSampler(local).start()
    time.sleep(0.3)
    la = eng.kernel_launches
    ms_dev, c0, c1 = timed(step_device, args.steps, args.warmup)
    launches_timed = timed.launches
    clock_windows = [(c0, c1)]
    value = world * nprob * iters * args.steps / (ms_dev * 1e-3)

    # verdicts of the last iterate (on the device), gathered as counts: the only cross-rank traffic
    d_ok = torch.empty(nprob, dtype=torch.int32, device=dev)
    d_first = torch.empty(nprob, dtype=torch.int32, device=dev)
    eng.verdict_device(nprob, d_g.data_ptr(), d_ok.data_ptr(), d_first.data_ptr())
    feas = d_ok.sum().to(torch.float64).reshape(1)
    if dist is not None:
        dist.all_reduce(feas)
    feasible_total = int(feas.item())

    # ---- roofline of k_constraints (the only kernel of the timed region) ----------------------------
    ln, un = eng.monomial_counts()
    NJ = eng.NJ
    reads = (int(ln.sum()) * (24 + 8) + int(un.sum()) * (8 + 8) + nprob * (T * NJ * 24 + T * NF * 8 + T * NJ * 144 +
                                                                         nobs * 96 + 56))
    writes = nprob * 8 * m * (1 + NF)
    alg_bytes = reads + writes
    n_launch = args.steps * iters
    kernel_ms = ms_dev / n_launch
    peak, peak_src = peaks()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"kernel": "k_constraints", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kernel_ms,
                "bytes_per_unit": alg_bytes / nprob}
    prof = os.path.join(ROOT, "profiles", "k_constraints_traffic.json")
    if os.path.exists(prof):  # dram bytes per launch from the committed ncu --set full capture of this command
        with open(prof) as f:
            pj = json.load(f)
        if pj.get("nprob") == nprob and pj.get("nobs") == nobs:
            roofline["traffic"] = pj.get("dram_bytes_per_launch")

    # ---- e2e: the host-buffer C-ABI call, copies inside the timed region --------------------------------
    e2e = None
    if not args.no_e2e:
        h_k = torch.empty((iters, nprob, NF), dtype=torch.float64, pin_memory=True)
        h_k.copy_(torch.from_numpy(ks_host))
        h_g = torch.empty((nprob, m), dtype=torch.float64, pin_memory=True)
        h_j = torch.empty((nprob, m, NF), dtype=torch.float64, pin_memory=True)
        k_np, g_np, j_np = h_k.numpy(), h_g.numpy(), h_j.numpy()

        def step_host():
            for it in range(iters):
                eng.eval_into(k_np[it], g_np, j_np)

        ms_e2e, h0, h1 = timed(step_host, args.steps, args.warmup)
        clock_windows.append((h0, h1))
        e2e = {"value": world * nprob * iters * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": iters * nprob * NF * 8, "d2h_bytes_per_step": iters * nprob * m * (1 + NF) * 8,
               "ms_per_step": ms_e2e / args.steps}
        # the host path must deliver what the device path computed
        eng.eval_device(nprob, d_k[iters - 1].data_ptr(), d_g.data_ptr(), d_j.data_ptr())
        torch.cuda.synchronize()
        if not (torch.equal(d_g.cpu(), h_g) and torch.equal(d_j.cpu(), h_j)):
            raise SystemExit("bench.py: host-buffer and device-pointer evaluations disagree")
        del h_g, h_j

    # ---- M1: reach-set build + one evaluation --------------------------------------------------------
    m1 = None
    if not args.no_m1:
        def step_m1():
            eng.build_device(nprob, nobs, tq0.data_ptr(), tqd0.data_ptr(), tqdd0.data_ptr(), tobs.data_ptr())
            eng.eval_device(nprob, d_k[0].data_ptr(), d_g.data_ptr(), d_j.data_ptr())

        ms_m1, h0, h1 = timed(step_m1, 2, 1)
        clock_windows.append((h0, h1))

        def step_build():
            eng.build_device(nprob, nobs, tq0.data_ptr(), tqd0.data_ptr(), tqdd0.data_ptr(), tobs.data_ptr())

        ms_b, _, _ = timed(step_build, 2, 0)
        m1 = {"metric": "reach-set build + eval_g + eval_jac_g per problem (M1)", "value": world * nprob * 2 / (ms_m1 * 1e-3),
              "unit": "problems/s", "ms_per_batch": ms_m1 / 2, "build_ms_per_batch": ms_b / 2,
              "build_us_per_problem": 1e3 * ms_b / 2 / nprob}
        if rank == 0:
            # latency of ONE planning iteration (BASELINE config 1; target < 1 ms): saved world 016_006
            csv = os.path.join(ROOT, "tests", "golden", "worlds", "scene_016_006.csv")
            a0, a1, a2, _, aobs = worlds.config1_problem(csv)
            e1 = ReachSetEngine(max_problems=1, max_obstacles=aobs.shape[0], device=local)
            e1.set_stream(stream.cuda_stream)
            t0_, t1_, t2_, to_ = (torch.tensor(a, dtype=torch.float64, device=dev) for a in (a0, a1, a2, aobs))
            k1 = torch.zeros(NF, dtype=torch.float64, device=dev)
            g1 = torch.empty(e1.lib.armour_num_constraints(e1._h, aobs.shape[0]), dtype=torch.float64, device=dev)
            j1 = torch.empty((g1.numel(), NF), dtype=torch.float64, device=dev)
            lat_b, lat_e = [], []
            for rep in range(25):
                ea, eb, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                ea.record(stream)
                e1.build_device(1, aobs.shape[0], t0_.data_ptr(), t1_.data_ptr(), t2_.data_ptr(), to_.data_ptr())
                eb.record(stream)
                e1.eval_device(1, k1.data_ptr(), g1.data_ptr(), j1.data_ptr())
                ec.record(stream)
                torch.cuda.synchronize()
                if rep >= 5:
                    lat_b.append(ea.elapsed_time(eb))
                    lat_e.append(eb.elapsed_time(ec))
            m1["single_problem_latency_ms"] = {"build": float(np.median(lat_b)), "eval": float(np.median(lat_e)),
                                               "total": float(np.median(lat_b) + np.median(lat_e)),
                                               "world": "scene_016_006.csv", "target_ms": 1.0}
            e1.close()
        fp64 = eng.measure_fp64_peak()
        # F_build: FP64 flops of the coefficient products of one build (counted by the oracle on a sample problem,
        # ~83-89 MFLOP per problem at the default threshold)
        m1["fp64_peak_tflops_measured"] = fp64

    # ---- batched device solver: whole NLP loops on the device, only k_opt and verdicts come back --------
    solver = None
    if not args.no_m1 and rank == 0:
        try:
            eng.solve(q_des)  # warm-up: workspace allocation
            t0s = time.perf_counter()
            ksol, oksol, _, itsol = eng.solve(q_des)
            dts = time.perf_counter() - t0s
            solver = {"metric": "plans/s (armour_batch_solve: trust-region SQP per world on the device, host in/out)",
                      "value": nprob / dts, "seconds": dts, "feasible": int(oksol.sum()),
                      "iterations_mean": float(itsol.mean()), "iterations_max": int(itsol.max()),
                      "constraint_evals_per_s": float((2 * itsol + 2).sum() / dts)}
        except Exception as exc:  # an extra, never allowed to take the bench line down
            solver = {"error": repr(exc)}

    clocks.stop()
    clk = clocks.summary(clock_windows)

    # ---- CPU baseline: the oracle on the host cores, bounded sample --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        build_s, times = oracle_run(cores, nobs, iters, 1, 0, seed=20261017, seconds=args.cpu_seconds)
        cpu = {"value": cores * iters * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cores} worlds of the same generator (one per host thread) x {iters} k-iterates x "
                         f"{len(times)} repetitions ({sum(times):.1f} s); oracle = C++ restatement of the reference",
               "build_s_per_world_1thread": build_s, "builds_per_s_all_cores": cores / build_s,
               "reference_build": reference_build_timing(nobs, 20261017)}
        if m1 is not None:
            # FP64 roofline of the build kernel: flops counted by the oracle on one problem of this batch
            from oracle.pyoracle import OracleProblem
            ref = OracleProblem(max_obstacles=max(40, nobs)).build(q0[0], qd0[0], qdd0[0], obs[0], nthreads=cores)
            fl = ref.stats()["flops"]
            ach = fl * nprob / (m1["build_ms_per_batch"] * 1e-3) / 1e12
            m1["roofline_build"] = {"kernel": "k_reachsets", "bound": "fp64", "flops_per_problem": fl,
                                    "achieved": ach, "peak": m1["fp64_peak_tflops_measured"], "unit": "TFLOP/s",
                                    "frac": ach / m1["fp64_peak_tflops_measured"],
                                    "note": "2x counted flops (nominal + interval RNEA) in the oracle; the kernel runs "
                                            "RNEA once with two radius lanes"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Kinova Gen3 batched {nprob} random worlds x {nobs} obstacles per GPU, one eval_g + "
                                   f"eval_jac_g per world per k-iterate, {iters} k-iterates per step",
                       "time_intervals": T, "obstacles": nobs, "constraints_per_world": m,
                       "k_iterates_per_step": iters, "worlds_per_gpu": nprob, "parallelism": f"worlds sharded x{world}",
                       "l2": "outputs (g + dense Jacobian) and reach-set tables per launch exceed the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "m1": m1, "solver": solver,
            "gpu_launches": int(launches_timed), "clocks": clk,
            "feasible_worlds_last_iterate": feasible_total, "build_launches": int(la - launches0),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
