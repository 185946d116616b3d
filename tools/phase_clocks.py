#!/usr/bin/env python3
"""Developer probe: where a solve_kernel3 warp spends its cycles.  Builds a variant of the
library with -DOSC_PHASE_CLOCKS (sums of clock64() at phase boundaries, tools/_build/, not the
product), runs cold and warm steps and prints cycles per environment per phase.
usage: phase_clocks.py [preset config n_envs]"""
import ctypes as C
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
sys.path.insert(0, ROOT)
import numpy as np


def build():
    import __graft_entry__ as g
    out = os.path.join(ROOT, "tools", "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libosc_b200_phase.so")
    src = os.path.join(g.PKG, "csrc", "osc_b200.cu")
    deps = [src] + [os.path.join(g.PKG, "csrc", f) for f in ("osc_core.cuh", "osc_core3.cuh", "osc_condensed.cuh", "osc_warp.cuh")]
    if g._newer(lib, deps):
        subprocess.run(["nvcc", *g.NVCC_FLAGS, "-DOSC_PHASE_CLOCKS", "-o", lib, src], check=True)
    return lib


def main():
    if "--build-only" in sys.argv:
        print(build())
        return
    import osc_b200 as ob
    from osc_b200 import capi
    capi.LIB_PATH = build()
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    preset = args[0] if len(args) > 0 else "walter_sr_true_tumbling_mjjoint"
    config = args[1] if len(args) > 1 else "tumbling"
    N = int(args[2]) if len(args) > 2 else 16384
    spec = ob.load_preset(preset)
    L = capi.load()
    L.osc_debug_phase_clocks.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(3)]
    g = capi.BatchedOSC(spec, N)
    g.enable_timing(True)
    buf = (C.c_ulonglong * 64)()

    def report(tag):
        L.osc_debug_phase_clocks(buf, 1)
        t = np.array(list(buf), dtype=np.uint64).astype(np.int64)  # wrap-around differences
        d = lambda a, b: float(np.int64(t[a] - t[b])) / N
        r = g.results()
        kt = g.read_timing()
        rows = [("prepare (assemble, iterates)", d(1, 0)), ("  scaling record -> D, E", d(13, 0)),
                ("  scale P, Aeq", d(14, 13)), ("  bounds, cost, pyramid", d(15, 14)),
                ("  iterates, rest", d(1, 15)), ("fetch next (TMA issue)", d(2, 1)),
                ("set_rho + factor", d(3, 2)), ("  Kd build", d(9, 2)), ("  Kd_dv^-1 (sweep)", d(10, 9)),
                ("  W products", d(11, 10)), ("  S product", d(12, 11)), ("  S^-1 + register loads", d(3, 12)),
                ("iterations", d(5, 4)), ("residuals", d(6, 5)), ("termination / rho update", d(7, 6)),
                ("outputs", d(8, 16)), ("  1 / c", d(32, 16)), ("  solution, state stores", d(33, 32)),
                ("  result scalars", d(8, 33)),
                ("residuals: exchange + group shuffles", d(34, 5)), ("residuals: rows, columns", d(35, 34)),
                ("residuals: maxima", d(36, 35)), ("residuals: reduce", d(6, 36)),
                ("whole solve part", d(8, 2)), ("environment total (excl. wait)", d(8, 0))]
        print(f"--- {tag}: solve kernel {kt.solve_ms:.3f} ms, iters mean {r['iters'].mean():.1f}")
        for k, v in rows:
            print(f"  {k:34s} {v:10.0f} cycles/env")
        it = r["iters"].mean()
        print(f"  cycles per iteration               {d(5, 4) / it:10.0f}")
        if t[45] != 0:
            print("  fused build + equilibration kernel (cycles/env at its own occupancy):")
            for k, v in [("wait for the landing stage", d(41, 40)), ("r = bias - t", d(42, 41)),
                         ("J'WJ (DMMA)", d(43, 42)), ("H, f -> shared + bulk stores", d(44, 43)),
                         ("Ruiz passes", d(45, 44)), ("total", d(45, 40))]:
                print(f"    {k:32s} {v:10.0f}")
        if t[30] > 0:
            print(f"  SM clock while the kernel ran (clock64 / globaltimer of CTA 0): {1e3 * t[29] / t[30]:.0f} MHz "
                  f"over {t[30] / 1e3:.1f} us")

    if "--condensed" in sys.argv:
        def creport(tag):
            L.osc_debug_phase_clocks(buf, 1)
            t = np.array(list(buf), dtype=np.uint64).astype(np.int64)
            d = lambda a, b: float(np.int64(t[a] - t[b])) / N
            r = g.results(); kt = g.read_timing()
            print(f"--- condensed {tag}: kernel {kt.solve_ms:.3f} ms, iters mean {r['iters'].mean():.1f}")
            for k, v in [("condense (Cholesky, G, d0, v)", d(18, 17)), ("P' rows (all passes)", d(20, 19)),
                         ("Ruiz + assemble + warm start", d(22, 21)), ("set_rho + factor (K^-1)", d(24, 23)),
                         ("admm (iterations + checks)", d(25, 24)), ("  iterations only", d(28, 27)),
                         ("outputs", d(26, 25)), ("environment total", d(26, 17))]:
                print(f"  {k:34s} {v:10.0f} cycles/env")
            print(f"  cycles per iteration               {d(28, 27) / r['iters'].mean():10.0f}")
        g.upload(steps[0]); g.step_condensed(); g.sync(); L.osc_debug_phase_clocks(buf, 1)
        g.reset_condensed(); g.step_condensed(); g.sync(); creport("cold")
        for rep in range(2):
            g.upload(steps[1 + rep % 2]); g.step_condensed(); g.sync(); creport("warm")
        return
    g.setup(steps[0]); g.step_device(); g.sync(); L.osc_debug_phase_clocks(buf, 1)
    g.setup(steps[0]); g.step_device(); g.sync(); report("cold")
    for rep in range(2):
        g.upload(steps[1 + rep % 2]); g.step_device(); g.sync(); report("warm")


if __name__ == "__main__":
    main()
