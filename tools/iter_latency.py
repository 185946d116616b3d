#!/usr/bin/env python3
"""Developer probe: cycles per ADMM iteration of one warp, alone and under full occupancy
(fixed iteration budget: eps = 0, no termination checks, no rho adaptation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
import numpy as np
import osc_b200 as ob
from osc_b200 import capi

if os.environ.get("OSC_LIB"):  # a developer build of the library (tools/_build/)
    capi.LIB_PATH = os.environ["OSC_LIB"]
preset = sys.argv[1] if len(sys.argv) > 1 else "walter_sr_true_tumbling_mjjoint"
config = sys.argv[2] if len(sys.argv) > 2 else "tumbling"
spec = ob.load_preset(preset)
sizes = [int(a) for a in sys.argv[3:]] or [1, 4, 8, 148, 1184, 16384]
for n_envs in sizes:
    res = {}
    for K in (200, 1200):
        st = capi.default_settings(eps_abs=0.0, eps_rel=0.0, max_iter=K, check_termination=0,
                                   adaptive_rho=0)
        g = capi.BatchedOSC(spec, n_envs, st)
        inp = ob.synth.make_inputs(spec, n_envs, config, step=0)
        g.setup(inp)
        g.enable_timing(True)
        for _ in range(3):
            g.setup()
            g.step_device(); g.sync()
        t = g.read_timing()
        res[K] = t.solve_ms
        g.close()
    per_iter_us = (res[1200] - res[200]) / 1000.0 * 1e3
    print(f"n_envs {n_envs:6d}: solve {res[200]:.3f} ms @200 it, {res[1200]:.3f} ms @1200 it -> "
          f"{per_iter_us:.3f} us per iteration = {per_iter_us * 1965:.0f} cycles")
