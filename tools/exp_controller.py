"""Developer tool: device-resident throughput of the controller kernels (states per second) and the CPU sides beside them."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from armour_b200 import RobustController  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "golden", "robot_models", "kinova_without_gripper.txt")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
c = RobustController(MODEL, 0.03)
c.set_stream(st.cuda_stream)
rng = np.random.default_rng(0)
host = [rng.uniform(-np.pi, np.pi, (n, 7))] + [rng.uniform(-2, 2, (n, 7)) for _ in range(4)]
q, qd, qda, qdd, qdd_des = [torch.tensor(a, dtype=torch.float64, device=dev) for a in host]
out = [torch.empty((n, 7), dtype=torch.float64, device=dev) for _ in range(3)]
stat = torch.empty(n, dtype=torch.int32, device=dev)
Kr = np.full(7, 10.0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


t_int = timed(lambda: c.rnea_device(n, q.data_ptr(), qd.data_ptr(), qda.data_ptr(), qdd.data_ptr(), d_tau_lo=out[1].data_ptr(),
                                    d_tau_hi=out[2].data_ptr()))
t_nom = timed(lambda: c.rnea_device(n, q.data_ptr(), qd.data_ptr(), qda.data_ptr(), qdd.data_ptr(), d_tau=out[0].data_ptr()))
t_upd = timed(lambda: c.update_device(n, Kr, 1.0, 1e-2, 1e-10, q.data_ptr(), qd.data_ptr(), qda.data_ptr(), qdd.data_ptr(),
                                      qdd_des.data_ptr(), d_u=out[0].data_ptr(), d_status=stat.data_ptr()))
print(f"n={n}: interval pass {t_int:.3f} ms = {n / t_int / 1e3:.2f} M states/s | nominal pass {t_nom:.3f} ms = {n / t_nom / 1e3:.2f} M/s | "
      f"controller update {t_upd:.3f} ms = {n / t_upd / 1e3:.2f} M states/s")
m = min(n, 1 << 16)
t0 = time.perf_counter()
c.update(Kr, 1.0, 1e-2, 1e-10, *(a[:m] for a in host))
t1 = time.perf_counter()
c.update(Kr, 1.0, 1e-2, 1e-10, *(a[:m] for a in host))
t2 = time.perf_counter()
print(f"host-pointer update of {m} states: {1e3 * (t2 - t1):.2f} ms = {m / (t2 - t1) / 1e6:.2f} M states/s (first call {1e3 * (t1 - t0):.1f} ms)")
from oracle import pycontroller  # noqa: E402
for name, cls in (("oracle", pycontroller.OracleController),) + ((("reference", pycontroller.ReferenceController),)
                                                                 if pycontroller.reference_available() else ()):
    o = cls()
    k = 2000
    t0 = time.perf_counter()
    for i in range(k):
        o.update(Kr, 1.0, 1e-2, 1e-10, *(a[i] for a in host))
    dt = time.perf_counter() - t0
    print(f"{name} (CPU, 1 thread, one state per call through ctypes): {1e6 * dt / k:.1f} us per controller update = {k / dt:.0f} states/s")
