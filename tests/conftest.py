import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ref_or_restated(orc):
    """Oracle loop nests driven by the reference's own kernel when oracle/_ref is there."""
    have = orc.ref_lib() is not None

    class Ctx:
        def __enter__(self):
            if have:
                orc.use_reference_kernel(True)
            return have

        def __exit__(self, *a):
            if have:
                orc.use_reference_kernel(False)

    return Ctx


@pytest.fixture(scope="session")
def hostcheck():
    """TEST-ONLY host build of the generated recurrences (tests/hostcheck.cpp)."""
    bdir = os.path.join(ROOT, "tests", "_build")
    so = os.path.join(bdir, "libhostcheck.so")
    csrc = os.path.join(ROOT, "rchem_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "hostcheck.cpp"), os.path.join(csrc, "basis_model.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in ("eri_core.h", "pair_build.h", "basis_model.h", "boys_delta.h", "boys_delta_build.h")]
    deps.append(os.path.join(ROOT, "rchem_b200", "gen", "gen_eri.py"))
    gen = os.path.join(csrc, "gen", "eri_class_list.h")
    if not os.path.exists(gen) or os.path.getmtime(gen) < os.path.getmtime(deps[-1]):
        subprocess.run([sys.executable, os.path.join(ROOT, "rchem_b200", "gen", "gen_eri.py"),
                        "--outdir", os.path.join(csrc, "gen")], check=True)
    deps.append(gen)
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(bdir, exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared",
                        "-ffp-contract=off", "-w", "-o", so] + srcs, check=True)
    H = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    H.hostcheck_nshells.argtypes = [C.c_int, dp, ip, ip, dp, dp, dp, ip, ip]
    H.hostcheck_shell_quartet.argtypes = [C.c_int, dp, ip, ip, dp, dp, dp] + [C.c_int] * 5 + [dp]
    H.hostcheck_boys.argtypes = [C.c_int, C.c_int, C.c_double, dp]
    H.hostcheck_boys.restype = None
    H.hostcheck_ref_tables_ok.restype = C.c_int
    H.hostcheck_last_min_x.restype = C.c_double
    H.hostcheck_last_proved_far.restype = C.c_int
    H.hostcheck_far_threshold.argtypes = [C.c_int]
    H.hostcheck_far_threshold.restype = C.c_double
    H.hostcheck_last_prims_used.restype = C.c_int
    H.hostcheck_set_fuse.argtypes = [C.c_int]
    H.hostcheck_set_fuse.restype = None
    return H


@pytest.fixture(scope="session")
def rc():
    import rchem_b200

    return rchem_b200


@pytest.fixture(scope="session")
def geo():
    from rchem_b200 import geometry

    return geometry
