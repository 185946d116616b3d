#!/usr/bin/env python3
"""Developer timing probe (not the contract bench): cold + warm control steps on resident inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
import numpy as np
import osc_b200 as ob
from osc_b200 import capi

def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "walter_sr_true_tumbling_mjjoint"
    config = sys.argv[2] if len(sys.argv) > 2 else "tumbling"
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    if os.environ.get("OSC_LIB"):  # a developer build of the library (tools/_build/)
        capi.LIB_PATH = os.environ["OSC_LIB"]
    spec = ob.load_preset(preset)
    print("dfma peak TF/s", capi.measure_dfma_tflops(0))
    t0 = time.time()
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(3)]
    print("gen s", time.time() - t0)
    g = capi.BatchedOSC(spec, N)
    g.setup(steps[0])
    g.enable_timing(True)
    for rep in range(3):
        g.setup(steps[0])
        g.step_device(); g.sync()
        t = g.read_timing()
        r = g.results()
        print(f"cold: build {t.build_ms:.3f} ms scale {t.scale_ms:.3f} ms solve {t.solve_ms:.3f} ms -> {N/(t.build_ms+t.scale_ms+t.solve_ms)*1e3:.3e} solves/s; iters mean {r['iters'].mean():.1f} hist {np.bincount(r['iters']//25)}")
    for rep in range(6):
        g.upload(steps[1 + rep % 2])
        g.step_device(); g.sync()
        t = g.read_timing()
        r = g.results()
        print(f"warm: build {t.build_ms:.3f} ms scale {t.scale_ms:.3f} ms solve {t.solve_ms:.3f} ms -> {N/(t.build_ms+t.scale_ms+t.solve_ms)*1e3:.3e} solves/s; iters mean {r['iters'].mean():.1f} status ok {(r['status']==1).mean():.4f}")
    b = spec.algorithmic_bytes
    print("alg bytes/solve", b, "build-kernel GB/s (J+bias+targets in, H+f out):", N * 8 * (spec.s*spec.nv + 2*spec.s + spec.nv*spec.nv + spec.nv) / (t.build_ms*1e-3) / 1e9)

if __name__ == "__main__":
    main()
