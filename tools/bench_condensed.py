#!/usr/bin/env python3
"""Condensed fast mode vs the reference-parity path, device-resident warm steps (developer tool).
usage: bench_condensed.py [preset config n_envs]   (env OSC_B200_COND_WARPS = 8 | 10 | 12 | 15)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
import numpy as np
import torch
import osc_b200 as ob
from osc_b200 import capi

preset = sys.argv[1] if len(sys.argv) > 1 else "walter_sr_true_tumbling_mjjoint"
config = sys.argv[2] if len(sys.argv) > 2 else "tumbling"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
spec = ob.load_preset(preset)
F = ("M", "C", "J", "bias", "targets", "mask")
sets = []
for t in range(4):
    inp = ob.synth.make_inputs(spec, N, config, step=t)
    sets.append({k: torch.from_numpy(inp[k]).cuda() for k in F})
kw = {}
if os.environ.get('OSC_SCALING'): kw['scaling'] = int(os.environ['OSC_SCALING'])
g = capi.BatchedOSC(spec, N, capi.default_settings(**kw))
def bind(t):
    g.bind_device_inputs(*[sets[t % 4][k].data_ptr() for k in F])
for mode in ("reference-parity", "condensed"):
    bind(0)
    if mode == "condensed":
        g.reset_condensed(); step = g.step_condensed
    else:
        g.setup(); step = g.step_device
    step(); g.sync()
    cold = g.results()["iters"].mean()
    for t in range(1, 4):
        bind(t); step()
    g.enable_timing(True)
    for t in range(4, 24):
        bind(t); step()
    kt = g.read_timing(); g.enable_timing(False)
    r = g.results()
    tot = kt.build_ms + kt.scale_ms + kt.solve_ms
    print(f"{mode:17s} {preset} N={N} warps={os.environ.get('OSC_B200_COND_WARPS', 'default')}: build {kt.build_ms:.4f} "
          f"scale {kt.scale_ms:.4f} solve {kt.solve_ms:.4f} ms -> {N / tot / 1e3:.2f} M solves/s; "
          f"iters warm {r['iters'].mean():.2f} cold {cold:.1f} solved {(r['status'] == 1).mean():.4f}")
