"""Kernel-logic test of the reach-set construction kernel WITHOUT a GPU.

tests/emu/emu_k1.cpp compiles armour_b200/csrc/k1_reachsets.cuh (the very source nvcc compiles for sm_100a)
with g++ against tests/emu/cuda_emu.h, a fiber emulator of the CUDA execution model (one ucontext fiber per
CUDA thread, __syncthreads / warp collectives yield to a scheduler).  A few (problem, interval) units are
built that way and compared with the oracle: identical k-only monomial key sets, coefficients within
1e-12, radii containing the oracle's and within 1e-10 relative.  This is test infrastructure: the product
library has no CPU path (tests/test_abi.py::test_no_gpu_means_loud_failure).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, WORLDS

NF, T = 7, 128
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libemu_k1.so")


@pytest.fixture(scope="module")
def emu(built):
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_k1.cpp", "cuda_emu.h")]
    srcs += [os.path.join(ROOT, "armour_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "armour_b200", "csrc"))
             if f.endswith((".cuh", ".h"))]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", EMU_LIB,
                        os.path.join(EMU_DIR, "emu_k1.cpp")], check=True, capture_output=True)
    lib = C.CDLL(EMU_LIB)
    dp, ip, sp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ushort)
    lib.emu_k1_build.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.c_double, C.c_double, dp, dp, dp, ip, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, sp, dp, ip, dp, dp, sp, dp, dp, dp, ip]
    return lib


def run_units(lib, q0, qd0, qdd0, ts, model_id=0, thr=5e-4, capL=64, capU=128, arena_words=6144,
              tab_s_bytes=49152, mass_unc=-1.0, inertia_unc=-1.0):
    NJ = 7 if model_id == 0 else 8
    kr = np.full(NF, np.pi / 48)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    q0, qd0, qdd0 = f(q0), f(qd0), f(qdd0)
    units = np.asarray(ts, dtype=np.int32)
    out = dict(nl=np.zeros(T * NJ, np.int32), cl=np.zeros((T * NJ, 3)), hl=np.zeros((T * NJ, capL), np.uint16),
               gl=np.zeros((T * NJ, capL, 3)), nu=np.zeros(T * NF, np.int32), cu=np.zeros(T * NF), ru=np.zeros(T * NF),
               hu=np.zeros((T * NF, capU), np.uint16), gu=np.zeros((T * NF, capU)), torque_radius=np.zeros((NF, T)),
               link_gens=np.zeros((T, NJ, 18)))
    stats = np.zeros(8, np.int32)
    dp, ip, sp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_ushort)
    d = lambda a: a.ctypes.data_as(dp)
    rc = lib.emu_k1_build(model_id, T, thr, d(kr), mass_unc, inertia_unc, d(q0), d(qd0), d(qdd0),
                          units.ctypes.data_as(ip), len(units), capL, capU, arena_words, tab_s_bytes,
                          out["nl"].ctypes.data_as(ip), d(out["cl"]), out["hl"].ctypes.data_as(sp), d(out["gl"]),
                          out["nu"].ctypes.data_as(ip), d(out["cu"]), d(out["ru"]), out["hu"].ctypes.data_as(sp),
                          d(out["gu"]), d(out["torque_radius"]), d(out["link_gens"]), stats.ctypes.data_as(ip))
    out["stats"] = stats
    return rc, out


def compare_units(out, ref_tables, ts, NJ=7):
    r = ref_tables
    for t in ts:
        for l in range(NJ):
            i = t * NJ + l
            n = r["nl"][i]
            assert out["nl"][i] == n, (t, l)
            assert np.array_equal(out["hl"][i, :n].astype(np.uint64), r["hl"][i, :n])
            if n:
                assert np.max(np.abs(out["gl"][i, :n] - r["gl"][i, :n])) <= 1e-12
            assert np.max(np.abs(out["cl"][i] - r["cl"][i])) <= 1e-12
            G, Gr = out["link_gens"][t, l].reshape(6, 3), r["link_gens"][t, l].reshape(6, 3)
            assert np.max(np.abs(G[:3] - Gr[:3])) <= 1e-12
            assert np.all(G[3:] >= Gr[3:])
            nz = Gr[3:] > 0
            assert np.max((G[3:][nz] - Gr[3:][nz]) / Gr[3:][nz]) <= 1e-10
        for j in range(NF):
            i = t * NF + j
            n = r["nu"][i]
            assert out["nu"][i] == n, (t, j)
            assert np.array_equal(out["hu"][i, :n].astype(np.uint64), r["hu"][i, :n])
            if n:
                assert np.max(np.abs(out["gu"][i, :n] - r["gu"][i, :n])) <= 1e-12
            assert abs(out["cu"][i] - r["cu"][i]) <= 1e-11
            assert out["ru"][i] >= r["ru"][i] and (out["ru"][i] - r["ru"][i]) / r["ru"][i] <= 1e-10
            a, b = out["torque_radius"][j, t], r["torque_radius"][j, t]
            assert a >= b and (a - b) / b <= 1e-10


def test_emulated_kernel_matches_oracle_saved_world(emu):
    from armour_b200 import worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, _, obs = worlds.config1_problem(os.path.join(WORLDS, "scene_016_006.csv"))
    ts = [0, 63, 127]
    rc, out = run_units(emu, q0, qd0, qdd0, ts)
    assert rc == 0 and out["stats"][2] == 0 and out["stats"][3] == len(ts)
    ref = OracleProblem().build(q0, qd0, qdd0, obs)
    compare_units(out, ref.tables(64, 128), ts)


def test_emulated_kernel_matches_oracle_moving_start(emu):
    from armour_b200 import worlds
    from oracle.pyoracle import OracleProblem
    q0, qd0, qdd0, _, obs = worlds.random_problems(1, 3, seed=5)
    ts = [1, 96]
    rc, out = run_units(emu, q0[0], qd0[0], qdd0[0], ts)
    assert rc == 0 and out["stats"][2] == 0
    ref = OracleProblem().build(q0[0], qd0[0], qdd0[0], obs[0])
    compare_units(out, ref.tables(64, 128), ts)


def test_small_shared_memory_spills_to_global_with_identical_results(emu):
    """With a tiny shared-memory arena / table pool the kernel continues in its global spill space
    (never truncates): bit-identical tables."""
    from armour_b200 import worlds
    q0, qd0, qdd0, _, _ = worlds.random_problems(1, 3, seed=9)
    rc_a, a = run_units(emu, q0[0], qd0[0], qdd0[0], [100])
    rc_b, b = run_units(emu, q0[0], qd0[0], qdd0[0], [100], arena_words=1024, tab_s_bytes=8192)
    assert rc_a == 0 and rc_b == 0
    assert b["stats"][1] > 0, "expected hash tables in the global pool"
    for key in ("nl", "cl", "hl", "gl", "nu", "cu", "ru", "hu", "gu", "torque_radius", "link_gens"):
        assert np.array_equal(a[key], b[key]), key


def test_mg_task_list_is_a_topological_order(emu):
    """The latency configuration builds one unit as a list of tasks claimed in a fixed order by three groups; a
    group waits for the inputs of the task it claimed, so the order must be topological (no deadlock) and hold every
    task exactly once.  The dependencies below are the mailbox imports of run_task (k1_reachsets.cuh)."""
    KINDS = ["W", "WA", "WD", "T4", "LA", "T10", "TF", "TN", "FKC", "FKL", "FB", "NB", "U", "EPI"]
    emu.emu_mg_task_list.argtypes = [C.c_int, C.POINTER(C.c_ushort)]
    emu.emu_mg_task_list.restype = C.c_int
    for NJ in (7, 8):
        buf = (C.c_ushort * 256)()
        n = emu.emu_mg_task_list(NJ, buf)
        tasks = [(KINDS[buf[i] >> 8], buf[i] & 255) for i in range(n)]
        assert len(tasks) == len(set(tasks)) == 13 * NJ + 1
        pos = {t: i for i, t in enumerate(tasks)}

        def deps(kind, i):
            prev = i - 1
            d = {
                "W": [("W", prev)], "WA": [("WA", prev)], "WD": [("WD", prev), ("WA", i)],
                "T4": [("WA", prev), ("W", prev)], "LA": [("LA", prev), ("WD", prev), ("T4", i)],
                "T10": [("WA", i), ("W", i)], "TF": [("WD", i), ("LA", i), ("T10", i)],
                "TN": [("WD", i), ("W", i), ("WA", i)], "FKC": [("FKC", prev)], "FKL": [("FKC", i)],
                "FB": [("FB", i + 1), ("TF", i)], "NB": [("NB", i + 1), ("TN", i), ("TF", i), ("FB", i)],
                "U": [("NB", i)], "EPI": [("U", j) for j in range(NJ)],
            }[kind]
            return [(k, j) for k, j in d if 0 <= j < NJ]

        for t in tasks:
            for dep in deps(*t):
                assert pos[dep] < pos[t], (NJ, t, dep)
