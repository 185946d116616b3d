#!/usr/bin/env python3
"""Batched roll-out that never leaves the GPU between control steps.

The reference's example drivers (examples/standing.cc:120-165,
examples/walter_sr_true_tumbling_mjjoint.cc:523-1019) do, per robot and per control step:
    site poses / velocities  -> task-space PD targets      (host loop)
    MuJoCo contacts          -> contact mask                (host loop)
    update_state / update_taskspace_targets / get_torque_command
Here N robots do all three on the device: `targets_pd`, `contact_mask_from_contacts` and
`step_device` read and write HBM only; the simulator that would produce the OSCData and the
site states is replaced by synthetic tensors (the MJCF models are not part of this repository).

    python examples/rollout_device_resident.py [n_envs] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "operational-space-control_b200", "python"))

import numpy as np
import torch

import osc_b200 as ob
from osc_b200 import capi


def main():
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    spec = ob.load_preset("walter_sr_true_tumbling_mjjoint")
    ns, nc = spec.ns, spec.nc
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)

    # OSCData of the batch (what update_osc_data would produce), resident in HBM
    inp = ob.synth.make_inputs(spec, n_envs, "tumbling", step=0)
    data = {k: torch.from_numpy(inp[k]).to(dev) for k in ("M", "C", "J", "bias")}
    osc = capi.BatchedOSC(spec, n_envs)
    # the handle reads M, C, J, bias from our tensors; targets and mask stay its own buffers,
    # which the two device-side helpers below fill
    osc.bind_device_inputs(M=data["M"].data_ptr(), C_=data["C"].data_ptr(),
                           J=data["J"].data_ptr(), bias=data["bias"].data_ptr())

    def unit(q):
        return q / q.norm(dim=-1, keepdim=True)

    sites = dict(
        pos=torch.randn(n_envs, ns, 3, dtype=torch.float64, device=dev, generator=g) * 0.05,
        quat=unit(torch.randn(n_envs, ns, 4, dtype=torch.float64, device=dev, generator=g)),
        vel=torch.zeros(n_envs, ns, 3, dtype=torch.float64, device=dev),
        angvel=torch.zeros(n_envs, ns, 3, dtype=torch.float64, device=dev),
        pos_des=torch.zeros(n_envs, ns, 3, dtype=torch.float64, device=dev),
        quat_des=unit(torch.randn(n_envs, ns, 4, dtype=torch.float64, device=dev, generator=g)),
    )
    kp_lin = np.full(ns, 150.0); kd_lin = np.full(ns, 25.0)      # examples/standing.cc:153-154
    kp_ang = np.full(ns, 50.0); kd_ang = np.full(ns, 10.0)
    wheel_geoms = np.array([3, 4, 7, 8, 11, 12, 15, 16], np.int32)  # wheel_sites_mujoco (:436)
    max_con = 16
    pairs = torch.randint(0, 20, (n_envs, max_con, 2), dtype=torch.int32, device=dev, generator=g)
    ncon = torch.randint(0, max_con + 1, (n_envs,), dtype=torch.int32, device=dev, generator=g)
    ptrs = {k: v.data_ptr() for k, v in sites.items()}
    torch.cuda.synchronize()

    def make_targets_and_mask():
        osc.targets_pd(ptrs, kp_lin, kd_lin, kp_ang, kd_ang)
        osc.contact_mask_from_contacts(pairs.data_ptr(), ncon.data_ptr(), max_con, wheel_geoms)

    make_targets_and_mask()
    osc.setup()                                    # initialize_optimization for every robot
    buf = osc.device_buffers()
    torque = torch.empty(n_envs, spec.nu, dtype=torch.float64, device=dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        # "simulate": the sites drift towards their targets (stands in for mj_step)
        sites["pos"].mul_(0.98)
        make_targets_and_mask()
        osc.step_device()                          # torques land in buf.torque (HBM)
    t1.record()
    torch.cuda.synchronize()
    import ctypes as C
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(torque.data_ptr()), C.c_void_p(buf.torque),
                                         C.c_size_t(torque.numel() * 8), 3)
    r = osc.results()
    ms = t0.elapsed_time(t1) / steps
    print(f"{n_envs} robots x {steps} control steps, all inputs and outputs in HBM: "
          f"{ms:.3f} ms per step = {n_envs / ms * 1e3:.3e} solves/s; "
          f"solved {float((r['status'] == 1).mean()):.3f}, mean iterations {r['iters'].mean():.1f}, "
          f"|tau|max {float(torque.abs().max()):.1f} N m")
    assert torch.isfinite(torque).all()


if __name__ == "__main__":
    main()
