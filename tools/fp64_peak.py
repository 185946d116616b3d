#!/usr/bin/env python3
"""FP64-FMA roofline denominator with its clock record (profiles/fp64_peak.json).

MEASURED_PEAKS.json (driver-written) has HBM and bf16 entries only, so the FP64 peak the solver
kernels are quoted against is measured here: osc_measure_dfma_tflops (dfma_peak_kernel: 8
independent DFMA chains per thread, 8 CTAs of 256 threads per SM, best of 4 timed launches),
while nvidia-smi samples the SM clock.  Nominal: 148 SMs x 64 FP64 lanes x 2 x 1.965 GHz = 37.2.
usage: fp64_peak.py [out.json]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
from osc_b200 import capi

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "fp64_peak.json")
f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,"
                      "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                      "clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-lms", "20", "-i", "0"],
                     stdout=f, stderr=subprocess.DEVNULL)
time.sleep(0.3)
runs = [capi.measure_dfma_tflops(0) for _ in range(8)]
time.sleep(0.1)
p.terminate(); p.wait(timeout=5)
f.flush(); f.seek(0)
rows = [[c.strip() for c in l.split(",")] for l in f.read().splitlines() if l.strip()]
sm = [float(r[0].split()[0]) for r in rows]
line = {"dfma_tflops_runs": runs, "dfma_tflops": max(runs),
        "nominal_tflops": 148 * 64 * 2 * 1.965e9 / 1e12,
        "how": "osc_measure_dfma_tflops: 8 independent DFMA chains per thread, 148 x 8 CTAs of 256 "
               "threads, 65536 iterations, best of 4 timed launches per run",
        "clocks": {"samples": len(sm), "sm_mhz_max_seen": max(sm) if sm else None,
                   "sm_mhz_median_upper_half": sorted(sm)[len(sm) * 3 // 4] if sm else None,
                   "sm_max_mhz": float(rows[0][1].split()[0]) if rows else None,
                   "power_w_max": max(float(r[2].split()[0]) for r in rows) if rows else None,
                   "any_slowdown": any("Active" in " ".join(r[3:5]) and "Not" not in r[3] for r in rows)},
        "note": "tcgen05 / TMEM have no FP64 kind, so the FP64 work of this path runs on the DFMA "
                "pipe and mma.sync m8n8k4 (DMMA); both share this peak on B200"}
json.dump(line, open(out, "w"), indent=1)
print(json.dumps(line))
os.unlink(f.name)
