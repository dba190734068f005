"""bench.py's clock sampler: which nvidia-smi samples enter the `clocks` object of the bench line (host logic, no GPU)."""
import importlib.util
import os

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _line(sm, reasons=("Not Active",) * 4):
    return f"0, {sm}, 1965, 400.0, 0x0, " + ", ".join(reasons)


def test_device_window_preferred_when_it_holds_enough_samples():
    s = _bench().ClockSampler(0)
    s.lines = [(10.00 + 0.05 * i, _line(1965)) for i in range(6)] + [(20.0, _line(1200))]
    out = s.summary([(10.0, 10.3), (19.9, 20.1)])
    assert out["samples"] == 6 and out["sm_mhz"] == 1965.0 and out["window"] == "device-timed region"


def test_short_device_window_falls_back_to_all_timed_regions_and_keeps_reasons():
    s = _bench().ClockSampler(0)
    hot = ("Not Active", "Not Active", "Active", "Not Active")
    s.lines = [(10.01, _line(1965)), (20.0, _line(1950)), (20.05, _line(1800, hot)), (30.0, _line(300))]
    out = s.summary([(10.0, 10.04), (19.9, 20.1)])
    assert out["samples"] == 3 and out["window"].startswith("all timed regions")
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_thermal_slowdown"]


def test_no_samples_is_reported_as_such():
    out = _bench().ClockSampler(0).summary([(0.0, 1.0)])
    assert out["samples"] == 0 and out["sm_mhz"] is None


def test_reference_arm_never_maps_the_product_library():
    """bench.py --impl reference must time the CPU side only: it may load oracle/ libraries, never libarmour_b200.so
    (round-1 verdict: the arm imported the armour_b200 package for its input generators and mapped the product)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, argparse; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--iters','1'];"
        "import bench; bench.main();"
        "maps=open('/proc/self/maps').read();"
        "assert 'armour_b200' not in sys.modules, 'package imported';"
        "assert 'libarmour_b200' not in maps, 'product library mapped';"
        "assert 'liboracle' in maps or 'libarmour_ref' in maps; print('CLEAN')")
    res = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "CLEAN" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
