#!/usr/bin/env python3
"""bench.py -- OSC solves/sec of the batched operational-space controller.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one control step of every environment: the hot path
(update_optimization_data -> update_optimization -> solve_optimization -> torque slice,
reference walter_sr/operational_space_controller.h:515-631) over one batch of synthetic
OSCData, warm-started from the previous step like the reference's control_loop.  Step t
uses the t-th "control tick" of the synthetic environments (1 % drift per tick), cycling
through NSETS resident batches.

value      whole-job solves/s with inputs resident in HBM (device-timed, max over ranks)
e2e        same metric through the C-ABI with pinned HOST buffers: H2D of the step's
           inputs + step + D2H of the torques inside the timed region
roofline   dominant kernel (solve_kernel3 for the Walter robots: assembly, factorisation,
           ADMM): algorithmic FLOPs / CUDA-event time vs the measured FP64-FMA peak;
           `roofline_scale` is the fused objective-build + equilibration kernel
           (build_scale_kernel3; FP64 pipe: one MAC per entry of J'WJ, one mul + compare per
           matrix entry and Ruiz pass); `roofline_build` the HBM-bound stand-alone build kernel
           (build_qp_kernel), timed in the three-kernel form of the step (`three_kernel_form`)
cpu_baseline  the oracle (restatement of the reference's CPU path, "port") on the box's
           host cores, bounded sample, same protocol
--impl reference   times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))
# stdout carries exactly one JSON line.  NCCL (and anything else in native code) prints to
# fd 1, so fd 1 is pointed at stderr for the whole run and the JSON line goes to a duplicate
# of the original stdout: NCCL's INFO lines (rank counts, NVLS / P2P transports) stay visible
# on stderr instead of being silenced.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
os.environ["NCCL_DEBUG"] = os.environ.get("OSC_BENCH_NCCL_DEBUG", "INFO")
os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2]/[4]: Walter Sr tumbling with body targets, 16384 envs per GPU
    "walter_sr_tumbling_16384_per_gpu": dict(preset="walter_sr_true_tumbling_mjjoint",
                                             config="tumbling", envs=16384),
    # configs[1]
    "go2_standing_4096": dict(preset="unitree_go2", config="go2_standing", envs=4096),
    # configs[3]
    "walter_sr_wheels_stairs_8192_per_gpu": dict(preset="walter_sr_wheels", config="stairs",
                                                 envs=8192),
    # configs[0] (the reference's own single-robot case, batched)
    "walter_sr_standing_4096": dict(preset="walter_sr", config="standing", envs=4096),
}
DEFAULT_WORKLOAD = "walter_sr_tumbling_16384_per_gpu"
NSETS = 4
FIELDS = ("M", "C", "J", "bias", "targets", "mask")
METRIC = "OSC solves/sec (whole box)"
UNIT = "solves/s"


def algorithmic_flops_per_solve(spec, iters, check_every=25):
    """dense-reduced-form operation count of the solver part (the H/f build
    is the other kernel): reduced system + Cholesky once, per-iteration solve + products,
    per-check residuals every 25 iterations.  `iters` = per-environment iteration counts:
    the number of residual checks is the mean of ceil(k / 25) per environment (not ceil of
    the mean, which counted two checks per solve for a mean of 25.1)."""
    n, nv, nc = spec.n, spec.nv, spec.nc
    iters = np.asarray(iters, dtype=np.float64)
    setup = nv * n * n + n ** 3 / 3.0
    nnzA = nv * n + 12 * nc + n
    per_iter = 2 * 2 * nnzA + 2 * n * n + 12 * spec.m
    per_check = 2 * n * n + 2 * nnzA
    return float(setup + per_iter * iters.mean() + per_check * np.ceil(iters / check_every).mean())


def condensed_flops_per_solve(spec, iters, check_every=25):
    """Algorithmic operation count of the condensed fast mode (n' = nu + 3 nc variables):
    Cholesky of M, G = M^-1 [B Jc | -C], Hd G, the symmetric P' = G' Hd G, an n' x n' SPD
    inverse, then per iteration one n' x n' mat-vec + friction rows + vector updates and per
    check P'x through G and Hd."""
    nv, nc = spec.nv, spec.nc
    n1 = spec.nu + 3 * nc
    m1 = 4 * nc + n1
    iters = np.asarray(iters, dtype=np.float64)
    setup = nv ** 3 / 3.0 + 2 * nv * nv * (n1 + 1) + 2 * nv * nv * n1 + nv * n1 * (n1 + 1) + n1 ** 3
    scaling = 10 * 2 * (n1 * n1 + 2 * (12 * nc + n1))
    per_iter = 2 * n1 * n1 + 2 * 2 * 12 * nc + 12 * m1
    per_check = 2 * (2 * nv * n1) + 2 * nv * nv + 2 * 2 * 12 * nc
    return float(setup + scaling + per_iter * iters.mean() + per_check * np.ceil(iters / check_every).mean())


def scale_ops_per_solve(spec, passes=10):
    """Equilibration kernel (OSQP scale_data): every pass takes the infinity norm of every
    column and row of the scaled [P A'; A 0], i.e. one multiply and one compare per stored
    entry of P and two of each per entry of A (row and column norm), all on the FP64 pipe."""
    n, nv, nu, nc = spec.n, spec.nv, spec.nu, spec.nc
    nnz_p = nv * nv + (n - nv)
    nnz_a = nv * (nv + 3 * nc) + nu + 12 * nc + n
    return passes * 2 * (nnz_p + 2 * nnz_a)


def build_macs_per_solve(spec):
    """Objective build: lower triangle of J'WJ (rows with a weight) and f = 2 J'W(bias - t)."""
    return spec.s * (spec.nv * (spec.nv + 1) // 2 + spec.nv)


def build_bytes_per_solve(spec):
    """build kernel: reads J, bias, targets; writes H (dv block) and f."""
    return 8 * (spec.s * spec.nv + 2 * spec.s + spec.nv * spec.nv + spec.nv)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv",
                                       "-lms", "20", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        lines = [l.strip() for l in self.f.read().splitlines()[1:] if l.strip()]
        sm, mx, reasons = [], [], set()
        for l in lines:
            c = [x.strip() for x in l.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1].split()[0]))
                mx.append(float(c[2].split()[0]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            top = sorted(sm)[len(sm) // 2:]  # samples under load are the upper half
            out.update(sm_mhz=float(np.median(top)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        return out


def bind_to_gpu_numa(local):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs of the NUMA node its
    GPU hangs off, so that the H2D stream does not cross the socket interconnect.  Returns
    (info dict, original affinity) -- the affinity is restored before the CPU baseline runs."""
    orig = None
    info = {"node": None, "cpus": None}
    try:
        orig = os.sched_getaffinity(0)
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader",
                              "-i", str(local)], capture_output=True, text=True, timeout=20).stdout
        bus = out.strip().splitlines()[0].strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]  # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= orig
            if cpus:
                os.sched_setaffinity(0, cpus)
                info = {"node": node, "cpus": len(cpus)}
    except Exception as e:  # no sysfs / nvidia-smi: keep the default placement
        info["error"] = type(e).__name__
    return info, orig


def device_peak_dfma(capi, local, cache={}):
    if local not in cache:
        cache[local] = capi.measure_dfma_tflops(local)
    return cache[local]


def measure_resident(ob, capi, sharding, torch, spec, wl, n_envs, steps, warmup, nsets, rank,
                     world, local, dev, hbm_peak, with_dual=False, mode="reference"):
    """Device-resident warm-step throughput of one (robot, config, batch size) point: the same
    protocol as the headline (inputs in HBM, `nsets` resident control ticks cycling, CUDA
    events on the launch stream, max over ranks), plus the per-kernel roofline fractions."""
    stream = torch.cuda.current_stream().cuda_stream
    dev_sets = []
    for t in range(nsets):
        inp = ob.synth.make_inputs(spec, n_envs, wl["config"], first_env=rank * n_envs, step=t)
        dev_sets.append({k: torch.from_numpy(inp[k]).to(dev) for k in FIELDS})
        del inp
    in_bytes = sum(v.numel() * 8 for v in dev_sets[0].values())
    osc = capi.BatchedOSC(spec, n_envs, device=local)

    def bind(t):
        d = dev_sets[t % nsets]
        osc.bind_device_inputs(*[d[k].data_ptr() for k in FIELDS])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    bind(0)
    if mode == "condensed":
        osc.reset_condensed(stream)
        step = osc.step_condensed
    else:
        osc.setup(stream=stream)
        step = osc.step_device
    for t in range(warmup):
        bind(t)
        step(stream)
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i, t in enumerate(range(warmup, warmup + steps)):
        bind(t)
        step(stream)
        evs[i + 1].record()
    barrier()
    ms = sharding.max_over_ranks(evs[0].elapsed_time(evs[-1]), dev) / steps
    p50 = sharding.max_over_ranks(
        float(np.median([a.elapsed_time(b) for a, b in zip(evs, evs[1:])])), dev)
    res = osc.results(stream)
    osc.enable_timing(True)  # per-kernel durations: a second loop (see main)
    for t in range(warmup + steps, warmup + 2 * steps):
        bind(t)
        step(stream)
    barrier()
    kt = osc.read_timing()
    osc.enable_timing(False)
    dfma = device_peak_dfma(capi, local)
    if mode == "condensed":
        flops = condensed_flops_per_solve(spec, res["iters"]) * n_envs
        return {"robot": spec.robot, "preset": wl["preset"], "synthetic_config": wl["config"],
                "envs_per_gpu": n_envs, "total_envs": world * n_envs, "steps": steps,
                "value": world * n_envs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                "iters_mean": float(res["iters"].mean()),
                "solved_frac": float((res["status"] == capi.SOLVED).mean()),
                "kernel_ms": {"build_qp_kernel": kt.build_ms, "condensed_kernel": kt.solve_ms},
                "roofline": {"kernel": "condensed_kernel", "bound": "fp64_fma",
                             "achieved": flops / (kt.solve_ms * 1e-3) / 1e12, "peak": dfma,
                             "unit": "TFLOP/s", "frac": flops / (kt.solve_ms * 1e-3) / 1e12 / dfma,
                             "algorithmic_flops_per_solve": flops / n_envs}}
    flops = algorithmic_flops_per_solve(spec, res["iters"]) * n_envs
    out = {"robot": spec.robot, "preset": wl["preset"], "synthetic_config": wl["config"],
           "envs_per_gpu": n_envs, "total_envs": world * n_envs, "steps": steps,
           "value": world * n_envs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "p50_batch_latency_ms": p50, "iters_mean": float(res["iters"].mean()),
           "solved_frac": float((res["status"] == capi.SOLVED).mean()),
           "inputs_exceed_l2": bool(nsets * in_bytes > 126e6),
           "kernel_ms": {"build_scale_kernel3": kt.scale_ms, "solve_kernel3": kt.solve_ms},
           "roofline": {"kernel": "solve_kernel3", "bound": "fp64_fma",
                        "achieved": flops / (kt.solve_ms * 1e-3) / 1e12, "peak": dfma,
                        "unit": "TFLOP/s",
                        "frac": flops / (kt.solve_ms * 1e-3) / 1e12 / dfma},
           "roofline_scale_frac": (scale_ops_per_solve(spec, 10) + build_macs_per_solve(spec))
                                  * n_envs / (kt.scale_ms * 1e-3) / 1e12 / (dfma / 2.0)}
    if with_dual:
        # contacts whose friction pyramid is active at the solution: a friction row with a
        # positive multiplier (rows nv .. nv + 4 nc of the dual) on an unmasked contact
        yf = res["y"][:, spec.nv:spec.nv + 4 * spec.nc].reshape(n_envs, spec.nc, 4)
        mk = dev_sets[(warmup + steps - 1) % nsets]["mask"].cpu().numpy() > 0
        act = (yf > 1e-6).any(2) & mk
        out["friction_cone_active_frac_envs"] = float(act.any(1).mean())
        out["friction_cone_active_frac_contacts"] = float(act.sum() / max(1, mk.sum()))
    osc.close()
    del dev_sets
    torch.cuda.empty_cache()
    return out


def one_robot_latency(ob, capi, spec, wl, local, reps=300):
    """BASELINE configs[0]: ONE Walter Sr in the reference's own loop -- host buffers in,
    torque out per control step (osc_step_host on a one-environment handle), wall clock."""
    steps = [ob.synth.make_inputs(spec, 1, wl["config"], step=t) for t in range(NSETS)]
    osc = capi.BatchedOSC(spec, 1, device=local)
    osc.setup(steps[0])
    tq = capi.pinned_empty((1, spec.nu))
    pin = []
    for st in steps:
        d = {}
        for k in FIELDS:
            a = capi.pinned_empty(st[k].shape)
            a[...] = st[k]
            d[k] = a
        pin.append(d)
    ptrs = [[d[k].ctypes.data for k in FIELDS] for d in pin]
    for t in range(20):
        osc.step_host_into(ptrs[t % NSETS], tq)
    lat = []
    for t in range(reps):
        t0 = time.perf_counter()
        osc.step_host_into(ptrs[t % NSETS], tq)
        lat.append(time.perf_counter() - t0)
    r = osc.results()
    osc.close()
    lat = np.array(lat) * 1e6
    return {"robot": spec.robot, "preset": wl["preset"], "envs": 1, "reps": reps,
            "p50_step_latency_us": float(np.median(lat)), "p99_step_latency_us": float(np.percentile(lat, 99)),
            "value": 1e6 / float(np.median(lat)), "unit": UNIT,
            "iters": int(r["iters"][0]), "note": "wall clock per control step: pinned host buffers "
            "in, 3 kernels, torque back to the host (osc_step_host, few-robot path)"}


def e2e_device_kinematics(ob, capi, sharding, torch, spec, wl, n_envs, steps, warmup, rank, world,
                          local, dev):
    """End to end with the MuJoCo-derived record made ON the device (osc_kinematics, SURVEY.md
    8f rank 2): per step the host hands over qpos, qvel, targets and the contact mask from
    pinned memory (0.4 kB per environment instead of the 13.4 kB OSCData record), the device
    computes M, C, J, bias for a SYNTHETIC tree of the robot's topology (the real MJCF models
    are external to the reference), runs the control step and returns the torques.  A
    different input distribution than the headline workload (M, J follow from the tree, not
    from the N(0, 0.3) generator): reported beside `e2e`, never instead of it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))  # model builder only (host-side constants)
    import osc_kinematics as okin
    tree = okin.walter_like_tree(0) if spec.nv == 14 else okin.go2_like_tree(0)
    model = capi.kin_model(tree)
    stream = torch.cuda.current_stream().cuda_stream
    sets = []
    for t in range(NSETS):
        qpos, qvel = okin.random_state(tree, n_envs, seed=t, first_env=rank * n_envs)
        inp = ob.synth.make_inputs(spec, 256, wl["config"], step=t)  # targets / mask pattern
        reps = (n_envs + 255) // 256
        tg = np.tile(inp["targets"], (reps, 1, 1))[:n_envs] * 0.02  # scaled to the tree's masses
        mk = np.tile(inp["mask"], (reps, 1))[:n_envs]
        pin = {}
        for k, a in (("qpos", qpos), ("qvel", qvel), ("targets", tg), ("mask", mk)):
            p = capi.pinned_empty(a.shape)
            p[...] = a
            pin[k] = p
        sets.append(pin)
    dq = torch.empty((n_envs, tree.nq), dtype=torch.float64, device=dev)
    dv = torch.empty((n_envs, tree.nv), dtype=torch.float64, device=dev)
    osc = capi.BatchedOSC(spec, n_envs, device=local)
    buf = osc.device_buffers()
    tq = capi.pinned_empty((n_envs, spec.nu))
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    H2D, D2H = 1, 2

    def copy(dst, src, nbytes, kind):
        rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dst), ctypes.c_void_p(src), ctypes.c_size_t(nbytes),
                                kind, ctypes.c_void_p(stream))
        assert rc == 0, rc

    def one(t, first=False):
        s = sets[t % NSETS]
        copy(dq.data_ptr(), s["qpos"].ctypes.data, s["qpos"].nbytes, H2D)
        copy(dv.data_ptr(), s["qvel"].ctypes.data, s["qvel"].nbytes, H2D)
        copy(buf.targets, s["targets"].ctypes.data, s["targets"].nbytes, H2D)
        copy(buf.mask, s["mask"].ctypes.data, s["mask"].nbytes, H2D)
        osc.kinematics(model, dq.data_ptr(), dv.data_ptr(), stream)
        if first:
            osc.setup(stream=stream)
        osc.step_device(stream)
        copy(tq.ctypes.data, buf.torque, tq.nbytes, D2H)
        osc.sync(stream)

    one(0, first=True)
    for t in range(warmup):
        one(t)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for t in range(warmup, warmup + steps):
        one(t)
    g1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = sharding.max_over_ranks(g0.elapsed_time(g1), dev) / steps
    r = osc.results(stream)
    h2d = sum(sets[0][k].nbytes for k in ("qpos", "qvel", "targets", "mask"))
    out = {"value": world * n_envs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": tq.nbytes,
           "iters_mean": float(r["iters"].mean()),
           "solved_frac": float((r["status"] == capi.SOLVED).mean()),
           "torques_finite": bool(np.isfinite(tq).all()),
           "inputs": "qpos, qvel, targets, mask from pinned host memory; M, C, J, bias computed on "
                     "the device for a synthetic tree of the robot's topology (osc_kinematics)"}
    osc.close()
    return out


def cpu_leg(spec, wl, sample_envs, steps, warmup, n_threads=0):
    """The reference's CPU path (oracle port), all host threads, same warm-step protocol on
    a bounded sample of the workload.  Returns (solves/s, info dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import osc_oracle as orc
    import osc_b200 as ob
    try:
        orc.build(native=True)
        native = True
    except Exception:
        native = False
    s = orc.default_settings(linsys=1)  # reduced Cholesky: the faster of the two equivalent forms
    b = orc.OracleBatch(spec, sample_envs, s, native=native)
    threads = n_threads or b.max_threads
    sets = [ob.synth.make_inputs(spec, sample_envs, wl["config"], step=t) for t in range(NSETS)]
    b.setup(sets[0], threads)
    iters = []
    for t in range(warmup):
        o = b.step(sets[t % NSETS], threads)
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        o = b.step(sets[t % NSETS], threads)
        iters.append(float(o["iters"].mean()))
    dt = time.perf_counter() - t0
    value = sample_envs * steps / dt
    info = dict(value=value, unit=UNIT, cores=threads, kind="port",
                sample=f"{sample_envs} envs x {steps} warm control steps of the same workload "
                       f"(oracle: restatement of the reference's QP build + OSQP 0.6.3, "
                       f"{'-march=native' if native else 'x86-64-v3'}, reduced-Cholesky KKT); "
                       f"{dt:.2f} s wall",
                iters_mean=float(np.mean(iters)), us_per_solve_per_core=1e6 * threads / value)
    return value, info


def run_reference(args, wl, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(wl["envs"], 2048)
    value, info = cpu_leg(spec, wl, sample, args.steps, args.warmup)
    ms = 1e3 * sample / value
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "envs_per_step": sample,
                       "warm_start": True, "note": "CPU path on host cores; each step is a "
                       "bounded sample of the workload"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the other BASELINE configs / sweep points (headline only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    import osc_b200 as ob
    wl = dict(WORKLOADS[args.workload])
    if args.envs_per_gpu:
        wl["envs"] = args.envs_per_gpu
    spec = ob.load_preset(wl["preset"])
    if args.impl == "reference":
        return run_reference(args, wl, spec)

    import torch
    import torch.distributed as dist
    from osc_b200 import capi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa, orig_affinity = bind_to_gpu_numa(local)  # before the pinned buffers are touched
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n_envs = wl["envs"]
    first_env = rank * n_envs  # weak scaling: fixed envs per GPU, disjoint environments
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic control ticks: NSETS batches in pinned host memory and in HBM
    host_sets, dev_sets = [], []
    for t in range(NSETS):
        inp = ob.synth.make_inputs(spec, n_envs, wl["config"], first_env=first_env, step=t)
        pinned = {}
        for k in FIELDS:
            a = capi.pinned_empty(inp[k].shape)
            a[...] = inp[k]
            pinned[k] = a
        host_sets.append(pinned)
        dev_sets.append({k: torch.from_numpy(inp[k]).to(dev) for k in FIELDS})
    in_bytes = sum(host_sets[0][k].nbytes for k in FIELDS)
    out_bytes = n_envs * spec.nu * 8

    osc = capi.BatchedOSC(spec, n_envs, device=local)

    def bind(t):
        d = dev_sets[t % NSETS]
        osc.bind_device_inputs(*[d[k].data_ptr() for k in FIELDS])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (value)
    bind(0)
    osc.setup(stream=stream)
    for t in range(args.warmup):
        bind(t)
        osc.step_device(stream)
    launches0 = osc.kernel_launches
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.3)  # let nvidia-smi start sampling before the timed region
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record()
    for i, t in enumerate(range(args.warmup, args.warmup + args.steps)):
        bind(t)
        osc.step_device(stream)
        evs[i + 1].record()
    barrier()
    ms_total = sharding.max_over_ranks(evs[0].elapsed_time(evs[-1]), dev)
    p50_ms = sharding.max_over_ranks(
        float(np.median([a.elapsed_time(b) for a, b in zip(evs, evs[1:])])), dev)
    clocks = sampler.stop() if sampler else None
    launches = osc.kernel_launches - launches0
    res = osc.results(stream)
    # per-kernel durations: a second, shorter loop with CUDA events around the kernels (the
    # four event records per step cost ~10 us, so they stay out of the loop `value` is timed on)
    osc.enable_timing(True)
    for t in range(args.warmup + args.steps, args.warmup + args.steps + max(10, min(args.steps, 50))):
        bind(t)
        osc.step_device(stream)
    barrier()
    kt = osc.read_timing()
    osc.enable_timing(False)
    ms_per_step = ms_total / args.steps
    value = world * n_envs / (ms_per_step * 1e-3)
    k_mean = float(res["iters"].mean())
    stats = sharding.reduce_stats(dict(solved=int((res["status"] == capi.SOLVED).sum()),
                                       iters=float(res["iters"].sum()),
                                       launches=int(launches)), world, dev)

    # ---- cold first step (reported beside the steady state)
    bind(0)
    osc.setup(stream=stream)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    osc.step_device(stream)
    c1.record()
    barrier()
    cold_ms = sharding.max_over_ranks(c0.elapsed_time(c1), dev)
    cold_iters = float(osc.results(stream)["iters"].mean())

    # ---- the three-kernel form of the step (stand-alone build_qp_kernel + scale_kernel3 +
    #      solve_kernel3; same results): what the fusion of the objective build into the
    #      equilibration kernel buys, and the HBM roofline of the build kernel on its own
    osc.set_fused_build(False)
    for t in range(args.warmup):
        bind(t)
        osc.step_device(stream)
    osc.enable_timing(True)
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n3 = max(10, min(args.steps, 50))
    u0.record()
    for t in range(args.warmup, args.warmup + n3):
        bind(t)
        osc.step_device(stream)
    u1.record()
    barrier()
    ms3 = sharding.max_over_ranks(u0.elapsed_time(u1), dev) / n3
    kt3 = osc.read_timing()
    osc.enable_timing(False)
    osc.set_fused_build(True)

    # ---- end to end through the C-ABI with host buffers (e2e)
    osc.bind_device_inputs()  # back to the handle's own input buffers
    tq = capi.pinned_empty((n_envs, spec.nu))
    ptrs = [[host_sets[t][k].ctypes.data for k in FIELDS] for t in range(NSETS)]
    osc.setup(host_sets[0], stream)
    for t in range(args.warmup):
        osc.step_host_into(ptrs[t % NSETS], tq, stream)
    barrier()
    w0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for t in range(args.warmup, args.warmup + args.steps):
        osc.step_host_into(ptrs[t % NSETS], tq, stream)
    g1.record()
    barrier()
    e2e_ms_local = g0.elapsed_time(g1) / args.steps
    e2e_ms = sharding.max_over_ranks(e2e_ms_local, dev) / 1.0
    e2e_wall_ms = sharding.max_over_ranks(1e3 * (time.perf_counter() - w0), dev) / args.steps
    e2e_value = world * n_envs / (e2e_ms * 1e-3)
    # bytes the C-ABI actually moved per step (it skips Jacobian rows nothing reads)
    e2e_h2d, e2e_d2h = osc.host_traffic()

    # ---- the same end-to-end step with the task Jacobian handed over in FP32 (opt-in
    #      osc_step_host_j32: widened on the device, FP64 from there on) -- reported separately
    #      with what the rounding of J does to the answer on this workload
    j32_sets = []
    for t in range(NSETS):
        a = capi.pinned_empty(host_sets[t]["J"].shape, np.float32)
        a[...] = host_sets[t]["J"]
        j32_sets.append(a)
    ptrs32 = [list(p) for p in ptrs]
    for t in range(NSETS):
        ptrs32[t][FIELDS.index("J")] = j32_sets[t].ctypes.data
    tq64 = tq.copy()  # torques of the last FP64-transport step (tick warmup + steps - 1)
    tq32 = capi.pinned_empty((n_envs, spec.nu))
    osc.setup(host_sets[0], stream)
    for t in range(args.warmup):
        osc.step_host_j32_into(ptrs32[t % NSETS], tq32, stream)
    barrier()
    j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    j0.record()
    for t in range(args.warmup, args.warmup + args.steps):
        osc.step_host_j32_into(ptrs32[t % NSETS], tq32, stream)
    j1.record()
    barrier()
    j32_ms = sharding.max_over_ranks(j0.elapsed_time(j1) / args.steps, dev)
    j32_h2d, _ = osc.host_traffic()
    dj = np.abs(tq32 - tq64)
    tolj = 1e-5 + 1e-4 * np.abs(tq64)
    e2e_j32 = {"value": world * n_envs / (j32_ms * 1e-3), "unit": UNIT, "ms_per_step": j32_ms,
               "h2d_bytes_per_step": j32_h2d, "d2h_bytes_per_step": e2e_d2h,
               "pcie_h2d_gbs": j32_h2d / (j32_ms * 1e-3) / 1e9,
               "torques_within_tol_of_fp64_transport": float((dj <= tolj).all(1).mean()),
               "worst_dtau_over_tol": float((dj / tolj).max()),
               "note": "opt-in osc_step_host_j32: J in FP32 over the host link, FP64 on the device; "
                       "same ticks as the e2e leg, compared after the same number of warm steps "
                       "(rank 0's environments)"}
    del j32_sets

    # ---- the all-gather of torques + statistics as peer stores behind the
    #      C-ABI (osc_gather_*): timed alone, and inside a second timed loop of resident
    #      steps so that `value_incl_gather` is measured, not derived
    sharding.setup_peer_gather(osc, rank, world, dev)
    osc.bind_device_inputs(*[dev_sets[0][k].data_ptr() for k in FIELDS])
    osc.setup(stream=stream)
    for t in range(args.warmup):
        bind(t)
        osc.step_device(stream)
        osc.gather_torques(stream)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for t in range(args.warmup, args.warmup + args.steps):
        bind(t)
        osc.step_device(stream)
        osc.gather_torques(stream)
    h1.record()
    barrier()
    ms_incl = sharding.max_over_ranks(h0.elapsed_time(h1), dev) / args.steps
    value_incl_gather = world * n_envs / (ms_incl * 1e-3)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(20):
        osc.gather_torques(stream)
    a1.record()
    barrier()
    gather_ms = sharding.max_over_ranks(a0.elapsed_time(a1), dev) / 20
    # cross-check of the peer-store path: every rank's gathered copy == an NCCL all-gather
    t_ptr, s_ptr = osc.gather_buffers()
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    got = torch.empty((world * n_envs, spec.nu), dtype=torch.float64, device=dev)
    rt.cudaMemcpy(ctypes.c_void_p(got.data_ptr()), ctypes.c_void_p(t_ptr),
                  ctypes.c_size_t(got.numel() * 8), 3)
    gstats = torch.empty((world, capi.GATHER_STATS), dtype=torch.float64, device=dev)
    rt.cudaMemcpy(ctypes.c_void_p(gstats.data_ptr()), ctypes.c_void_p(s_ptr),
                  ctypes.c_size_t(gstats.numel() * 8), 3)
    mine = torch.from_numpy(osc.torques(stream)).to(dev)
    gather_check, nccl_gather_ms = True, None
    if world > 1:
        ref = torch.empty_like(got)
        dist.all_gather_into_tensor(ref, mine)  # NCCL lazy init, untimed
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        dist.all_gather_into_tensor(ref, mine)
        n1.record()
        barrier()
        nccl_gather_ms = sharding.max_over_ranks(n0.elapsed_time(n1), dev)
        ok = torch.equal(ref, got)
        gather_check = bool(sharding.max_over_ranks(0.0 if ok else 1.0, dev) == 0.0)
    else:
        gather_check = bool(torch.equal(got, mine))
    gstats = gstats.cpu().numpy()
    # per-rank PCIe rate of the end-to-end leg (shows switch / socket sharing at N > 1)
    my_gbs = e2e_h2d / (e2e_ms_local * 1e-3) / 1e9
    if world > 1:
        allg = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(allg, torch.tensor([my_gbs], dtype=torch.float64, device=dev))
        pcie_per_rank = [float(x.item()) for x in allg]
        alln = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(alln, torch.tensor([float(numa["node"] if numa["node"] is not None else -1)],
                                           dtype=torch.float64, device=dev))
        numa_per_rank = [int(x.item()) for x in alln]
    else:
        pcie_per_rank = [my_gbs]
        numa_per_rank = [numa["node"] if numa["node"] is not None else -1]

    # ---- the other BASELINE configs and the per-GPU sweep points, same protocol (few steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    configs = None
    if not args.no_configs and args.workload == DEFAULT_WORKLOAD and not args.envs_per_gpu:
        osc.close()
        del dev_sets, host_sets
        torch.cuda.empty_cache()
        configs = {}
        ksteps = max(10, min(args.steps, 30))
        mr = lambda name, n, nsets, **kw: measure_resident(  # noqa: E731
            ob, capi, sharding, torch, ob.load_preset(WORKLOADS[name]["preset"]), WORKLOADS[name],
            n, ksteps, args.warmup, nsets, rank, world, local, dev, hbm_peak, **kw)
        configs["C2_go2_standing_4096"] = mr("go2_standing_4096", 4096, NSETS)
        configs["C4_walter_sr_wheels_stairs_8192_per_gpu"] = mr(
            "walter_sr_wheels_stairs_8192_per_gpu", 8192, NSETS, with_dual=True)
        sweep = []
        for n in (1024, 4096, 65536, 131072):  # 16384 per GPU is the headline itself
            sweep.append(mr(DEFAULT_WORKLOAD, n, 2 if n > 16384 else NSETS))
        configs["C5_walter_sr_sweep_per_gpu"] = sweep
        # the condensed fast mode (north_star subsystems (1)/(2)), reported separately: same
        # QP, same OSQP tolerances, different iterates -- never the reference-parity path
        fast = {"mode": "condensed (osc_step_condensed): Cholesky of M, G = M^-1 [B Jc], QP in "
                        "(u, z) only; parity = its own oracle (oracle/osc_condensed.py), "
                        "tests/test_gpu_parity.py::test_condensed_*"}
        for key, name, n in (("walter_sr_tumbling_16384_per_gpu", DEFAULT_WORKLOAD, 16384),
                             ("go2_standing_4096", "go2_standing_4096", 4096)):
            r = mr(name, n, NSETS, mode="condensed")
            fast[key] = r
        configs["fast_mode_condensed"] = fast
        configs["e2e_device_kinematics_walter_sr_16384_per_gpu"] = e2e_device_kinematics(
            ob, capi, sharding, torch, ob.load_preset(WORKLOADS[DEFAULT_WORKLOAD]["preset"]),
            WORKLOADS[DEFAULT_WORKLOAD], 16384, ksteps, args.warmup, rank, world, local, dev)
        if world == 1:
            wl1 = WORKLOADS["walter_sr_standing_4096"]
            configs["C1_walter_sr_standing_one_robot"] = one_robot_latency(
                ob, capi, ob.load_preset(wl1["preset"]), wl1, local)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    dfma_peak = capi.measure_dfma_tflops(local)
    flops = algorithmic_flops_per_solve(spec, res["iters"]) * n_envs
    solve_tflops = flops / (kt.solve_ms * 1e-3) / 1e12
    build_gbs = build_bytes_per_solve(spec) * n_envs / (kt3.build_ms * 1e-3) / 1e9
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    tr = traffic.get(args.workload, {})
    split = True  # equilibration runs in its own kernel (scale_kernel3) for every robot
    solve_name = "solve_kernel3"
    roofline = {"kernel": solve_name, "bound": "fp64_fma", "achieved": solve_tflops,
                "peak": dfma_peak, "unit": "TFLOP/s", "frac": solve_tflops / dfma_peak,
                "traffic": tr.get(solve_name + "_dram_bytes_per_launch"),
                "peak_source": "DFMA micro-benchmark run inside this bench "
                               "(MEASURED_PEAKS.json has no FP64 entry)",
                "algorithmic_flops_per_solve": flops / n_envs, "iters_mean": k_mean,
                "ncu": "latency-bound: 8 warps/SM at 255 registers (registers are granted per four "
                       "warps: the next step is 12 warps at 168), issue slots 39 %, FP64 pipe 34 %, "
                       "shared-memory pipe 57 % of peak (profiles/r3_ncu_summary.md)",
                "launch_ms": kt.solve_ms,
                "hbm_view": {"achieved_gbs": spec.algorithmic_bytes * n_envs
                             / ((kt.solve_ms + kt.scale_ms) * 1e-3) / 1e9,
                             "peak_gbs": hbm_peak}}
    roofline_scale = None
    if split:
        ops = (scale_ops_per_solve(spec, 10) + build_macs_per_solve(spec)) * n_envs
        t_ops = ops / (kt.scale_ms * 1e-3) / 1e12
        roofline_scale = {"kernel": "build_scale_kernel3", "bound": "fp64_pipe", "achieved": t_ops,
                          "peak": dfma_peak / 2.0,
                          "unit": "Tops/s (FP64 MAC of J'WJ / mul / compare)",
                          "frac": t_ops / (dfma_peak / 2.0),
                          "traffic": tr.get("build_scale_kernel3_dram_bytes_per_launch"),
                          "peak_source": "half the DFMA FLOP rate: one FP64-pipe instruction "
                                         "per lane per op (a DMMA m8n8k4 occupies the pipe like "
                                         "the 8 DFMAs per lane it replaces)",
                          "algorithmic_ops_per_solve": ops / n_envs,
                          "of_which_objective_build": build_macs_per_solve(spec),
                          "launch_ms": kt.scale_ms}
    roofline_build = {"kernel": "build_qp_kernel", "bound": "hbm", "achieved": build_gbs,
                      "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak,
                      "traffic": tr.get("build_qp_kernel_dram_bytes_per_launch"),
                      "peak_source": hbm_src, "launch_ms": kt3.build_ms,
                      "algorithmic_bytes_per_solve": build_bytes_per_solve(spec),
                      "note": "stand-alone build kernel, timed in the three-kernel form of the "
                              "step (osc_set_fused_build(0)); the default step runs the build "
                              "inside build_scale_kernel3"}
    three_kernel_form = {"ms_per_step": ms3, "value": world * n_envs / (ms3 * 1e-3), "steps": n3,
                         "kernel_ms": {"build_qp_kernel": kt3.build_ms,
                                       "scale_kernel3": kt3.scale_ms,
                                       "solve_kernel3": kt3.solve_ms}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "robot": spec.robot, "preset": wl["preset"],
                       "envs_per_gpu": n_envs, "total_envs": world * n_envs,
                       "warm_start": True, "osqp_settings": "OsqpSettings() defaults",
                       "inputs_exceed_l2": bool(in_bytes > 126e6),
                       "resident_input_sets": NSETS, "parallelism": f"env-sharded x{world}"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d,
                    "d2h_bytes_per_step": e2e_d2h, "ms_per_step": e2e_ms,
                    "wall_ms_per_step": e2e_wall_ms, "host_input_bytes_per_step": in_bytes,
                    "pcie_h2d_gbs": e2e_h2d / (e2e_ms * 1e-3) / 1e9,
                    "pcie_h2d_gbs_per_rank": pcie_per_rank, "numa_node_per_rank": numa_per_rank,
                    "limit": "PCIe: the OSCData record (M, J, bias, ...) is 13.4 kB per environment "
                             "and crosses the host link every step; at N > 1 ranks that share a "
                             "PCIe switch / socket share its bandwidth (see per-rank rates)"},
            "e2e_fp32_jacobian": e2e_j32,
            "gpu_launches": int(stats["launches"]),
            "roofline": roofline, "roofline_scale": roofline_scale,
            "roofline_build": roofline_build, "three_kernel_form": three_kernel_form,
            "kernel_ms": {"build_scale_kernel3": kt.scale_ms, solve_name: kt.solve_ms},
            "p50_batch_latency_ms": p50_ms,
            "solved_frac": stats["solved"] / (world * n_envs),
            "cold_start": {"ms_per_step": cold_ms, "value": world * n_envs / (cold_ms * 1e-3),
                           "iters_mean": cold_iters},
            "value_incl_gather": value_incl_gather, "ms_per_step_incl_gather": ms_incl,
            "torque_all_gather_ms": gather_ms,
            "gather": {"how": "peer stores over NVLink into CUDA-IPC-mapped slabs (osc_gather_torques,"
                              " one kernel, no collective call)", "ms": gather_ms,
                       "bytes_per_rank": n_envs * spec.nu * 8 * world, "check_vs_nccl": gather_check,
                       "nccl_all_gather_into_tensor_ms": nccl_gather_ms,
                       "stats_all": {k: [float(v) for v in gstats[:, i]]
                                     for i, k in enumerate(capi.GATHER_STAT_NAMES)}},
            "numa": numa}
    if configs is not None:
        line["configs"] = configs
    if world == 1 and not args.no_cpu_baseline:
        if orig_affinity:
            try:
                os.sched_setaffinity(0, orig_affinity)  # the CPU arm gets every host core
            except Exception:
                pass
        sample = min(n_envs, 2048)
        _, info = cpu_leg(spec, wl, sample, steps=40, warmup=2)
        line["cpu_baseline"] = info
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
