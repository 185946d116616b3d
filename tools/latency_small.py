#!/usr/bin/env python3
"""Developer probe: wall-clock latency of one control step through osc_step_host for small
batches (the reference's own use: one robot, 1 kHz loop).  usage: latency_small.py [preset config]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
import numpy as np
import osc_b200 as ob
from osc_b200 import capi

preset = sys.argv[1] if len(sys.argv) > 1 else "walter_sr"
config = sys.argv[2] if len(sys.argv) > 2 else "standing"
spec = ob.load_preset(preset)
for n in (1, 8, 64, 512):
    steps = [ob.synth.make_inputs(spec, n, config, step=t) for t in range(4)]
    g = capi.BatchedOSC(spec, n)
    g.setup(steps[0])
    tq = np.empty((n, spec.nu))
    ptrs = [g._ptrs(s) for s in steps]
    for k in range(20):
        g.step_host_into(ptrs[k % 4][1], tq)
    T = 300
    t0 = time.perf_counter()
    for k in range(T):
        g.step_host_into(ptrs[k % 4][1], tq)
    dt = (time.perf_counter() - t0) / T
    r = g.results()
    print(f"{preset} n_envs {n:4d}: {dt * 1e6:8.1f} us per step (host buffers in, torques out), iters mean {r['iters'].mean():.1f}")
