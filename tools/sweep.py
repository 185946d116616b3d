#!/usr/bin/env python3
"""BASELINE.json configs[4]: throughput sweep over the batch size on one GPU (device-resident
inputs, warm control steps, CUDA-event kernel times)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
import numpy as np
import osc_b200 as ob
from osc_b200 import capi

preset = sys.argv[1] if len(sys.argv) > 1 else "walter_sr_true_tumbling_mjjoint"
config = sys.argv[2] if len(sys.argv) > 2 else "tumbling"
sizes = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 64, 1024, 4096, 16384, 65536, 262144]
spec = ob.load_preset(preset)
rows = []
for N in sizes:
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(2)]
    g = capi.BatchedOSC(spec, N)
    g.setup(steps[0])
    g.step(steps[0])
    g.enable_timing(True)
    ms = []
    for rep in range(6):
        g.upload(steps[1 - rep % 2])
        g.step_device(); g.sync()
        t = g.read_timing()
        if rep >= 2:
            ms.append(t.build_ms + t.scale_ms + t.solve_ms)
    r = g.results()
    row = dict(n_envs=N, ms_per_step=float(np.median(ms)), solves_per_s=N / (np.median(ms) * 1e-3),
               iters_mean=float(r["iters"].mean()), solved=float((r["status"] == 1).mean()))
    rows.append(row)
    print(json.dumps(row), flush=True)
    g.close()
