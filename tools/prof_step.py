#!/usr/bin/env python3
"""Minimal driver for ncu: setup, one cold step, one warm step (resident inputs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
import osc_b200 as ob
from osc_b200 import capi

_a = [a for a in sys.argv[1:] if not a.startswith("--")]
preset = _a[0] if len(_a) > 0 else "walter_sr_true_tumbling_mjjoint"
config = _a[1] if len(_a) > 1 else "tumbling"
N = int(_a[2]) if len(_a) > 2 else 4096
spec = ob.load_preset(preset)
s0 = ob.synth.make_inputs(spec, N, config, step=0)
s1 = ob.synth.make_inputs(spec, N, config, step=1)
g = capi.BatchedOSC(spec, N)
g.setup(s0)
g.step_device(); g.sync()
g.upload(s1)
g.step_device(); g.sync()
r = g.results()
print("ok iters", r["iters"].mean(), "solved", (r["status"] == 1).mean())
if "--condensed" in sys.argv:  # the condensed fast mode: one cold and one warm step
    g.upload(s0); g.step_condensed(); g.sync()
    g.upload(s1); g.step_condensed(); g.sync()
    r = g.results()
    print("condensed ok iters", r["iters"].mean(), "solved", (r["status"] == 1).mean())
