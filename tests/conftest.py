import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure)."""
    import osc_oracle
    osc_oracle.build()
    return osc_oracle


@pytest.fixture(scope="session")
def host_core():
    """tests/host_core: the product's per-environment algorithm instantiated with one
    lane on the host -- a checker of the device code, never a product path."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "host_core")
    so = os.path.join(d, "libhost_core.so")
    srcs = [os.path.join(d, "host_core.cpp"),
            os.path.join(ROOT, "operational-space-control_b200", "csrc", "osc_core.cuh"),
            os.path.join(ROOT, "operational-space-control_b200", "csrc", "osc_core3.cuh"),
            os.path.join(ROOT, "operational-space-control_b200", "csrc", "osc_condensed.cuh"),
            os.path.join(ROOT, "operational-space-control_b200", "csrc", "osc_warp.cuh"),
            os.path.join(ROOT, "operational-space-control_b200", "csrc", "osc_params.h"),
            os.path.join(ROOT, "include", "osc_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                        "-o", so, srcs[0]], check=True)
    return C.CDLL(so)


@pytest.fixture(scope="session")
def host_core_reversed(tmp_path_factory):
    """The same harness with the emulated lanes of every body run in the opposite order."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "host_core")
    so = os.path.join(str(tmp_path_factory.mktemp("hc_rev")), "libhost_core_rev.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                    "-DOSC_WARP_REVERSE", "-o", so, os.path.join(d, "host_core.cpp")], check=True)
    return C.CDLL(so)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
