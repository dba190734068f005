"""The C-ABI library loads and exports every symbol include/armour_b200.h declares (no GPU needed)."""
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol(built):
    from armour_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "armour_b200.h")).read()
    declared = set(re.findall(r"\b(armour_[a-z_0-9]+)\s*\(", header))
    declared -= {"armour_config", "armour_ctx", "armour_reachset_tables"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in armour_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "python binding table out of sync with the header"
    assert lib.armour_abi_version() == 1


def test_config_default_and_status_strings(built):
    import ctypes as C
    from armour_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config()
    assert lib.armour_config_default(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(_lib.Config)
    assert cfg.num_time_steps == 128 and cfg.max_obstacles == 40
    assert abs(cfg.simplify_threshold - 5e-4) < 1e-20
    assert lib.armour_status_string(0) == b"ok"
    assert lib.armour_status_string(-3) == b"too many obstacles"
    assert lib.armour_config_default(None) == -1


def test_no_gpu_means_loud_failure(built):
    """Without a CUDA device context creation must fail with ARMOUR_ERR_CUDA, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from armour_b200 import ReachSetEngine, ArmourError
    with pytest.raises(ArmourError) as ei:
        ReachSetEngine()
    assert ei.value.code == -2


def test_product_never_touches_the_oracle():
    """Nothing under armour_b200/ or include/ may import, include or link oracle/ code."""
    bad = []
    for base in ("armour_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".so", ".pyc", ".o")):
                    continue
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle/|pyoracle|liboracle|from oracle|import oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
