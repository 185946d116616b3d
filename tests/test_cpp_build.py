"""CPU suite: the drop-in C++ headers meet a compiler in every configuration.

* the three robot headers (walter_sr, unitree_go2, walter_sr_wheels) x {OSCData injected,
  MuJoCo branch against tests/stubs/mujoco/mujoco.h + the scripted backend} compile and link
  against libosc_b200.so (the -m gpu suite runs them);
* when the reference tree is present (this container, not the GPU box): the reference's own
  example drivers -- UNMODIFIED, read where they lie under /root/reference/examples -- compile
  and link against the drop-in headers with stub Eigen / abseil / GLFW / runfiles / MuJoCo
  headers (tests/stubs/ref_example), i.e. the public surface the examples use
  (reference examples/walter_sr_standing.cc:89-167) is source compatible.
"""
import os
import subprocess

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "operational-space-control_b200")
STUBS = os.path.join(ROOT, "tests", "stubs")
REF_EXAMPLES = "/root/reference/examples"
LINK = ["-L", PKG, "-losc_b200", f"-Wl,-rpath,{PKG}", "-lpthread"]


@pytest.fixture(scope="module", autouse=True)
def _lib():
    if not os.path.exists(os.path.join(PKG, "libosc_b200.so")):
        import __graft_entry__ as g
        g.build()


@pytest.mark.parametrize("macro", ["", "-DROBOT_GO2", "-DROBOT_WW"])
def test_controller_headers_compile_and_link(tmp_path, macro):
    base = ["g++", "-std=c++20", "-O0", "-Wall", "-I", os.path.join(ROOT, "include")]
    if macro:
        base.append(macro)
    subprocess.run(base + [os.path.join(ROOT, "tests", "cpp", "test_controller.cpp"),
                           "-o", str(tmp_path / "a")] + LINK, check=True)
    subprocess.run(base + ["-I", STUBS, os.path.join(ROOT, "tests", "cpp", "test_controller_mujoco.cpp"),
                           os.path.join(STUBS, "fake_mujoco.cc"), "-o", str(tmp_path / "b")] + LINK,
                   check=True)


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference tree not present")
@pytest.mark.parametrize("example", ["walter_sr_standing.cc", "walter_sr_tumbling.cc",
                                     "standing.cc", "push_up.cc"])
def test_reference_examples_compile_unmodified(tmp_path, example):
    cmd = ["g++", "-std=c++20", "-O0", "-I", os.path.join(STUBS, "ref_example"), "-I", STUBS,
           "-I", os.path.join(ROOT, "include"), os.path.join(REF_EXAMPLES, example),
           os.path.join(STUBS, "fake_mujoco.cc"), os.path.join(STUBS, "ref_example", "viewer_stubs.cc"),
           "-o", str(tmp_path / "example")] + LINK
    subprocess.run(cmd, check=True)
