#!/usr/bin/env python3
"""Regenerate tests/golden/oracle_cases.npz: small seeded cases of every BASELINE config
with the oracle's results (setup + 3 control steps, OSQP default settings).

The fixture pins the oracle against drift (CPU test) and gives the GPU tests a
committed reference that does not depend on rebuilding the oracle on the GPU box.
Inputs are not stored: they are regenerated from (preset, config, seed) by
osc_b200.synth.make_inputs, whose determinism the CPU tests also check via a checksum.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import osc_b200 as ob  # noqa: E402
import osc_oracle as orc  # noqa: E402

CASES = [("walter_sr", "standing"), ("unitree_go2", "go2_standing"),
         ("walter_sr_true_tumbling_mjjoint", "tumbling"), ("walter_sr_wheels", "stairs")]
N_ENVS, N_STEPS = 32, 3


def checksum(inp):
    h = hashlib.sha256()
    for k in ("M", "C", "J", "bias", "targets", "mask"):
        h.update(np.ascontiguousarray(inp[k]).tobytes())
    return h.hexdigest()


def main():
    out = {}
    for preset, config in CASES:
        spec = ob.load_preset(preset)
        b = orc.OracleBatch(spec, N_ENVS, orc.default_settings())
        for t in range(N_STEPS):
            inp = ob.synth.make_inputs(spec, N_ENVS, config, step=t)
            if t == 0:
                assert b.setup(inp) == 0
            o = b.step(inp)
            key = f"{preset}|{config}|{t}"
            out[key + "|torque"] = o["torque"]
            out[key + "|iters"] = o["iters"]
            out[key + "|status"] = o["status"]
            out[key + "|rho"] = o["rho"]
            out[key + "|x"] = o["x"]
            out[key + "|y"] = o["y"]
            out[key + "|margin"] = o["margin"]
            out[key + "|sha"] = np.frombuffer(checksum(inp).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "oracle_cases.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
