"""Replay of vectors recorded from the REAL reference controller (tools/reference_dump).

A user who can build vannem95/operational-space-control (Bazel + MuJoCo + OSQP) records
OSCData / targets / mask -> solution / dual / exit code / iterations per control step with
tools/reference_dump/osc_dump.h and drops the converted file at tests/golden/reference_*.npz.
This module replays every such file -- Init, then each control step with the solver state
carried over, like control_loop (reference walter_sr/operational_space_controller.h:604-647)
-- through the oracle (CPU suite) and through the CUDA path (-m gpu) and compares, per step:
iterations and exit code bit-exactly, torques = solution[nv:nv+nu] within 1e-5 + 1e-4 |tau|.
No such file can be produced in this repository's container (none of the reference's
dependencies exist here), so without one the replay tests SKIP; the format itself --
recorder, converter, replayer -- is exercised on every run with the oracle playing the
reference (test_dump_format_round_trip).
"""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "reference_dump"))
import dump_to_npz  # noqa: E402

ATOL, RTOL = 1e-5, 1e-4
FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "reference_*.npz")))
# osqp-cpp OsqpExitCode (what the reference stores, :591) -> OSQP status value
EXIT_TO_STATUS = {0: 1, 1: -3, 2: -4, 3: 2, 4: 3, 5: 4, 6: -2, 9: -7}


def _inputs(g, prefix, t=None):
    keys = ("M", "C", "J", "bias", "targets", "mask")
    if t is None:
        return {k: g[prefix + k][None] for k in keys}
    return {k: g[k][t][None] for k in keys}


def replay(g, make, settings_for):
    """make(spec, settings) -> object with setup(inp) and step(inp) -> dict(torque, iters, status).
    Returns (interval used, number of steps)."""
    import osc_b200 as ob
    spec = ob.load_preset(str(g["preset"]))
    assert [spec.nv, spec.nu, spec.nc, spec.ns, spec.n, spec.m] == list(g["shape"])
    T = len(g["iterations"])
    given = int(g["adaptive_rho_interval"])
    candidates = [given] if given else [0, 25, 50, 75]
    report = []
    for interval in candidates:
        b = make(spec, settings_for(adaptive_rho_interval=interval))
        b.setup(_inputs(g, "init_"))
        bad = None
        for t in range(T):
            o = b.step(_inputs(g, "", t))
            tau = g["solution"][t][spec.nv:spec.nv + spec.nu]
            tol = ATOL + RTOL * np.abs(tau)
            want_status = EXIT_TO_STATUS.get(int(g["exit_code"][t]), None)
            if (int(o["iters"][0]) != int(g["iterations"][t]) or int(o["status"][0]) != want_status
                    or not (np.abs(o["torque"][0] - tau) <= tol).all()):
                bad = (t, int(o["iters"][0]), int(g["iterations"][t]), int(o["status"][0]),
                       want_status, float((np.abs(o["torque"][0] - tau) / tol).max()))
                break
        if bad is None:
            return interval, T
        report.append((interval, bad))
    raise AssertionError("no adaptive-rho interval reproduces the recording; first differing "
                         f"step per candidate (interval, (step, iters, want, status, want, ratio)): {report}")


class _OracleOne:
    def __init__(self, spec, settings):
        import osc_oracle as orc
        self.b = orc.OracleBatch(spec, 1, settings)

    def setup(self, inp):
        assert self.b.setup(inp) == 0

    def step(self, inp):
        return self.b.step(inp)


class _GpuOne:
    def __init__(self, spec, settings):
        from osc_b200 import capi
        self.g = capi.BatchedOSC(spec, 1, settings)

    def setup(self, inp):
        self.g.setup(inp)

    def step(self, inp):
        tq = self.g.step(inp)
        r = self.g.results()
        return dict(torque=tq, iters=r["iters"], status=r["status"])


@pytest.mark.parametrize("path", FILES or [None])
def test_oracle_replays_reference_vectors(oracle, path):
    if path is None:
        pytest.skip("no tests/golden/reference_*.npz (see tools/reference_dump/README.md)")
    interval, T = replay(np.load(path), _OracleOne, lambda **kw: oracle.default_settings(**kw))
    print(f"{os.path.basename(path)}: {T} control steps reproduced (adaptive_rho_interval {interval})")


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES or [None])
def test_gpu_replays_reference_vectors(path):
    if path is None:
        pytest.skip("no tests/golden/reference_*.npz (see tools/reference_dump/README.md)")
    from osc_b200 import capi
    interval, T = replay(np.load(path), _GpuOne, lambda **kw: capi.default_settings(**kw))
    print(f"{os.path.basename(path)}: {T} control steps reproduced on the GPU (interval {interval})")


RECORDER_MAIN = r'''
#include <cstdio>
#include <vector>
#include "osc_dump.h"
// reads a blob of records produced by the test and feeds them to the recorder the way the
// patched reference would (README.md): one record_init, then record_step per control step
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  int hdr[7];
  if (!f || fread(hdr, 4, 7, f) != 7) return 2;
  const int nv = hdr[0], nu = hdr[1], nc = hdr[2], ns = hdr[3], n = hdr[4], m = hdr[5], T = hdr[6];
  (void)nu;
  const size_t s = 6 * (size_t)ns, in = (size_t)nv * nv + nv + s * nv + 2 * s + nc;
  std::vector<double> a(in), x(n), y(m);
  auto& rec = osc_dump::Recorder::instance();
  rec.open({nv, hdr[1], nc, ns, n, m});
  if (!rec.active()) return 3;
  for (int t = -1; t < T; ++t) {
    if (fread(a.data(), 8, in, f) != in) return 2;
    const double *M = a.data(), *C = M + nv * nv, *J = C + nv, *b = J + s * nv, *tg = b + s, *mk = tg + s;
    if (t < 0) { rec.record_init(M, C, J, b, tg, mk); continue; }
    int tail[2];
    if (fread(x.data(), 8, n, f) != (size_t)n || fread(y.data(), 8, m, f) != (size_t)m ||
        fread(tail, 4, 2, f) != 2) return 2;
    rec.record_step(M, C, J, b, tg, mk, x.data(), y.data(), tail[0], tail[1]);
  }
  return 0;
}
'''


def test_dump_format_round_trip(oracle, tmp_path):
    """Recorder (C++, as it would be compiled into the reference) -> converter -> replayer, with
    the oracle in the role of the reference: a 6-step Walter Sr sequence, explicit interval."""
    import osc_b200 as ob
    spec = ob.load_preset("walter_sr")
    T, interval = 6, 50
    steps = [ob.synth.make_inputs(spec, 1, "tumbling", step=t) for t in range(T)]
    b = oracle.OracleBatch(spec, 1, oracle.default_settings(adaptive_rho_interval=interval))
    b.setup(steps[0])
    status_to_exit = {v: k for k, v in EXIT_TO_STATUS.items()}
    blob = tmp_path / "records.bin"
    keys = ("M", "C", "J", "bias", "targets", "mask")
    with open(blob, "wb") as fh:
        fh.write(np.array([spec.nv, spec.nu, spec.nc, spec.ns, spec.n, spec.m, T], np.int32).tobytes())
        for k in keys:
            fh.write(np.ascontiguousarray(steps[0][k][0]).tobytes())
        for t in range(T):
            o = b.step(steps[t])
            for k in keys:
                fh.write(np.ascontiguousarray(steps[t][k][0]).tobytes())
            fh.write(o["x"][0].tobytes())
            fh.write(o["y"][0].tobytes())
            fh.write(np.array([status_to_exit[int(o["status"][0])], int(o["iters"][0])], np.int32).tobytes())
    src = tmp_path / "rec.cpp"
    src.write_text(RECORDER_MAIN)
    exe = tmp_path / "rec"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "tools", "reference_dump"),
                    str(src), "-o", str(exe)], check=True)
    dump = tmp_path / "walter.oscdump"
    subprocess.run([str(exe), str(blob)], check=True, env=dict(os.environ, OSC_DUMP_FILE=str(dump)))
    shape, init, recs = dump_to_npz.read_dump(str(dump))
    assert len(recs) == T and shape["n"] == spec.n
    npz = tmp_path / "reference_walter_sr_selftest.npz"
    np.savez_compressed(npz, **dump_to_npz.to_npz(shape, init, recs, "walter_sr", interval))
    used, n = replay(np.load(npz), _OracleOne, lambda **kw: oracle.default_settings(**kw))
    assert (used, n) == (interval, T)
    # and the replayer does notice a wrong recording
    g = dict(np.load(npz))
    g["solution"] = g["solution"].copy()
    g["solution"][3, spec.nv] += 1.0
    with pytest.raises(AssertionError):
        replay(g, _OracleOne, lambda **kw: oracle.default_settings(**kw))


@pytest.mark.gpu
def test_gpu_replayer_on_an_oracle_recorded_sequence(oracle):
    """The GPU replayer (what test_gpu_replays_reference_vectors runs on a user's file) on a
    sequence recorded from the oracle, so that it is exercised without a reference recording."""
    import osc_b200 as ob
    spec = ob.load_preset("unitree_go2")
    T = 8
    steps = [ob.synth.make_inputs(spec, 1, "go2_standing", step=t) for t in range(T)]
    b = oracle.OracleBatch(spec, 1, oracle.default_settings())
    b.setup(steps[0])
    status_to_exit = {v: k for k, v in EXIT_TO_STATUS.items()}
    keys = ("M", "C", "J", "bias", "targets", "mask")
    outs = [b.step(s) for s in steps]
    g = {"preset": np.array("unitree_go2"), "adaptive_rho_interval": np.array(0),
         "shape": np.array([spec.nv, spec.nu, spec.nc, spec.ns, spec.n, spec.m], np.int32),
         "solution": np.stack([o["x"][0] for o in outs]), "dual": np.stack([o["y"][0] for o in outs]),
         "exit_code": np.array([status_to_exit[int(o["status"][0])] for o in outs], np.int32),
         "iterations": np.array([int(o["iters"][0]) for o in outs], np.int32)}
    for k in keys:
        g["init_" + k] = steps[0][k][0]
        g[k] = np.stack([s[k][0] for s in steps])
    from osc_b200 import capi
    used, n = replay(g, _GpuOne, lambda **kw: capi.default_settings(**kw))
    assert n == T
