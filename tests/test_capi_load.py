"""CPU tests of the drop-in boundary: the library builds, loads and exports every symbol include/lscqp.h declares;
the product fails loudly without a device (no CPU fallback); the oracle is not imported by the package."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "lscqp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lscqp_[a-z_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from lsc_dr_planner_b200 import capi
    lib = capi.load()
    names = _declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert set(capi.EXPORTS) <= set(names)
    assert b"sm_100a" in lib.lscqp_version()


def test_library_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "lsc_dr_planner_b200", "liblscqp.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_create_fails_loudly_without_device_or_on_bad_config():
    import torch
    from lsc_dr_planner_b200 import capi
    from lsc_dr_planner_b200 import workloads as W
    lib = capi.load()
    h = C.c_void_p()
    bad = capi.make_config(W.PlannerConfig(M=7))
    assert lib.lscqp_create(C.byref(bad), 0, C.byref(h)) == -1          # LSCQP_E_INVALID
    assert b"unsupported" in lib.lscqp_last_error()
    bad = capi.make_config(W.PlannerConfig(max_obs=4096))                        # above the ABI's obstacle capacity
    assert lib.lscqp_create(C.byref(bad), 0, C.byref(h)) == -1
    if not torch.cuda.is_available():
        ok = capi.make_config(W.PlannerConfig())
        assert lib.lscqp_create(C.byref(ok), 0, C.byref(h)) == -2       # LSCQP_E_NODEVICE: no CPU fallback
        with pytest.raises(capi.LscqpError):
            from lsc_dr_planner_b200.planner import BatchPlanner
            BatchPlanner(W.PlannerConfig())


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lsc_dr_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("(or, in tests, by the oracle)", ""), os.path.join(dirpath, f)
    code = "import sys; sys.path.insert(0, %r); import lsc_dr_planner_b200.capi, lsc_dr_planner_b200.planner, lsc_dr_planner_b200.workloads; " \
           "assert not any(m.startswith('oracle') for m in sys.modules)" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)


def test_jerk_gram_in_host_code_matches_oracle():
    """the product's own constant (csrc/host_common.hpp jerk_gram, through the emulator build) equals buildQBase"""
    import emul
    from lsc_dr_planner_b200 import workloads as W
    from oracle import oracle as orc
    Q = np.zeros(36)
    emul.lib().emul_jerk_gram.restype = None
    emul.lib().emul_jerk_gram(5, 3, C.c_double(0.2), Q.ctypes.data_as(C.c_void_p))
    assert np.allclose(Q.reshape(6, 6), orc.qbase(5, 3, 1, 0.2), rtol=1e-13, atol=1e-6)


def test_cpp_shim_compiles():
    """include/lscqp_shim.hpp (TrajOptimizer / CollisionConstraints / BatchTrajOptimizer) is valid C++17 on its own"""
    src = os.path.join(ROOT, "tests", "shim", "shim_smoke.cpp")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", src], check=True)


def test_cpp_shim_compiles_against_project_types():
    """... and in LSCQP_SHIM_EXTERNAL_TYPES mode, where Param / Mission / Agent / point3d / Trajectory come from the host
    project (the way it is used inside the reference tree): compiled against stand-ins with the reference's member lists"""
    src = os.path.join(ROOT, "tests", "shim", "shim_external.cpp")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", src], check=True)
