"""ctypes binding of the product library `libosc_b200.so` (C-ABI: include/osc_b200.h)
and `BatchedOSC`, the batched sibling of the reference's OperationalSpaceController.

Loading fails loudly when the library is missing; every call fails with the
library's error message when no B200 is present.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .specs import RobotSpec

_PKG = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
LIB_PATH = os.path.join(_PKG, "libosc_b200.so")

OSC_MAX_SITES = 32
OSC_MAX_NU = 16

SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED, UNSOLVED = 1, 2, -2, -10

EXPORTS = (
    "osc_default_settings", "osc_create", "osc_destroy", "osc_last_error", "osc_num_envs",
    "osc_get_device_buffers", "osc_upload", "osc_setup", "osc_step", "osc_reset_warm_start",
    "osc_download", "osc_sync", "osc_step_host", "osc_kernel_launches",
    "osc_measure_dfma_tflops", "osc_host_alloc", "osc_host_free", "osc_bind_device_inputs",
    "osc_timing_enable", "osc_timing_read", "osc_download_objective", "osc_reinit_count",
    "osc_targets_pd", "osc_contact_mask_from_contacts", "osc_host_traffic",
    "osc_selftest_warp",
    "osc_gather_create", "osc_gather_attach", "osc_gather_torques", "osc_gather_buffers",
    "osc_step_condensed", "osc_reset_condensed", "osc_kinematics",
    "osc_walter_tumbling_default_gains", "osc_targets_walter_tumbling",
    "osc_set_fused_build", "osc_step_host_j32",
)
KIN_MAX_BODIES = 16
IPC_HANDLE_BYTES = 64
GATHER_STATS = 8
GATHER_STAT_NAMES = ("n_envs", "solved", "iters_sum", "iters_max", "pri_res_max", "dua_res_max",
                     "reinits", "sequence")


class CRobotSpec(C.Structure):
    _fields_ = [
        ("nv", C.c_int), ("nu", C.c_int), ("nc", C.c_int), ("ns", C.c_int),
        ("w_trans", C.c_double * OSC_MAX_SITES), ("w_rot", C.c_double * OSC_MAX_SITES),
        ("w_torque", C.c_double), ("w_reg", C.c_double), ("mu", C.c_double),
        ("u_lb", C.c_double * OSC_MAX_NU), ("u_ub", C.c_double * OSC_MAX_NU),
        ("fz_max", C.c_double),
    ]


class CSettings(C.Structure):
    """Same field names as osqp::OsqpSettings (the subset that changes iterates)."""
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("adaptive_rho_tolerance", C.c_double),
        ("scaling", C.c_int), ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
        ("max_iter", C.c_int), ("check_termination", C.c_int), ("warm_start", C.c_int),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
    ]

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        # OsqpSettings() defaults for the two trailing fields when built positionally
        if len(args) < 13 and "eps_prim_inf" not in kw:
            self.eps_prim_inf = 1e-4
        if len(args) < 14 and "eps_dual_inf" not in kw:
            self.eps_dual_inf = 1e-4


class CDeviceBuffers(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "M", "C", "J", "bias", "targets", "mask", "torque", "solution", "dual", "iters", "status",
        "pri_res", "dua_res", "rho")]


class CSiteState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "pos", "quat", "vel", "angvel", "pos_des", "quat_des", "vel_des", "angvel_des")]


class CWalterTumblingGains(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("shin_kp", "shin_kv", "shin_rate", "thigh_kp",
                                          "thigh_kv", "thigh_rate", "thigh_height_offset")]


class CKinModel(C.Structure):
    _fields_ = [("nb", C.c_int), ("ns", C.c_int), ("parent", C.c_int * 16),
                ("bpos", (C.c_double * 3) * 16), ("bquat", (C.c_double * 4) * 16),
                ("jaxis", (C.c_double * 3) * 16), ("mass", C.c_double * 16),
                ("ipos", (C.c_double * 3) * 16), ("inertia", (C.c_double * 6) * 16),
                ("site_body", C.c_int * 32), ("site_pos", (C.c_double * 3) * 32),
                ("gravity", C.c_double * 3)]


def kin_model(tree, gravity=(0.0, 0.0, -9.81)) -> CKinModel:
    """osc_kin_model from a tree description with the attributes of oracle/osc_kinematics.Tree
    (parent, bpos, bquat, jaxis, mass, ipos, inertia [nb,3,3], site_body, site_pos)."""
    m = CKinModel()
    m.nb, m.ns = int(tree.nb), int(tree.ns)
    for b in range(tree.nb):
        m.parent[b] = int(tree.parent[b])
        m.mass[b] = float(tree.mass[b])
        for k in range(3):
            m.bpos[b][k] = float(tree.bpos[b][k])
            m.jaxis[b][k] = float(tree.jaxis[b][k])
            m.ipos[b][k] = float(tree.ipos[b][k])
        for k in range(4):
            m.bquat[b][k] = float(tree.bquat[b][k])
        I = tree.inertia[b]
        for k, v in enumerate((I[0][0], I[1][1], I[2][2], I[0][1], I[0][2], I[1][2])):
            m.inertia[b][k] = float(v)
    for s in range(tree.ns):
        m.site_body[s] = int(tree.site_body[s])
        for k in range(3):
            m.site_pos[s][k] = float(tree.site_pos[s][k])
    for k in range(3):
        m.gravity[k] = float(gravity[k])
    return m


class CKernelTimes(C.Structure):
    _fields_ = [("build_ms", C.c_float), ("solve_ms", C.c_float), ("steps", C.c_int),
                ("scale_ms", C.c_float)]


class OscError(RuntimeError):
    pass


_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise OscError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                       "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.osc_default_settings.argtypes = [C.POINTER(CSettings)]
    L.osc_create.argtypes = [C.POINTER(CRobotSpec), C.POINTER(CSettings), C.c_int, C.c_int,
                             C.POINTER(vp)]
    L.osc_destroy.argtypes = [vp]
    L.osc_last_error.argtypes = [vp]
    L.osc_last_error.restype = C.c_char_p
    L.osc_num_envs.argtypes = [vp]
    L.osc_get_device_buffers.argtypes = [vp, C.POINTER(CDeviceBuffers)]
    L.osc_upload.argtypes = [vp] + [vp] * 6 + [vp]
    L.osc_setup.argtypes = [vp, vp]
    L.osc_step.argtypes = [vp, vp]
    L.osc_reset_warm_start.argtypes = [vp, vp]
    L.osc_download.argtypes = [vp] + [vp] * 8 + [vp]
    L.osc_sync.argtypes = [vp, vp]
    L.osc_reinit_count.argtypes = [vp, ip, vp]
    L.osc_download_objective.argtypes = [vp, vp, vp, vp]
    L.osc_step_host.argtypes = [vp] + [vp] * 7 + [vp]
    L.osc_step_host_j32.argtypes = [vp] + [vp] * 7 + [vp]
    L.osc_set_fused_build.argtypes = [vp, C.c_int]
    L.osc_kernel_launches.argtypes = [vp]
    L.osc_kernel_launches.restype = C.c_longlong
    L.osc_measure_dfma_tflops.argtypes = [C.c_int, dp]
    L.osc_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.osc_host_free.argtypes = [vp]
    L.osc_bind_device_inputs.argtypes = [vp] + [vp] * 6
    L.osc_targets_pd.argtypes = [vp, C.POINTER(CSiteState), dp, dp, dp, dp, vp]
    L.osc_contact_mask_from_contacts.argtypes = [vp, vp, vp, C.c_int, ip, ip, vp]
    L.osc_selftest_warp.argtypes = [C.c_int, dp, dp]
    L.osc_host_traffic.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.osc_gather_create.argtypes = [vp, C.c_int, C.c_int, vp]
    L.osc_gather_attach.argtypes = [vp, vp, C.POINTER(vp)]
    L.osc_gather_torques.argtypes = [vp, vp]
    L.osc_gather_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.osc_kinematics.argtypes = [vp, C.POINTER(CKinModel), vp, vp, vp]
    L.osc_walter_tumbling_default_gains.argtypes = [C.POINTER(CWalterTumblingGains)]
    L.osc_targets_walter_tumbling.argtypes = [vp, C.POINTER(CWalterTumblingGains)] + [vp] * 6 + [
        C.c_double, C.c_double, vp]
    L.osc_step_condensed.argtypes = [vp, vp]
    L.osc_reset_condensed.argtypes = [vp, vp]
    L.osc_timing_enable.argtypes = [vp, C.c_int]
    L.osc_timing_read.argtypes = [vp, C.POINTER(CKernelTimes)]
    _LIB = L
    return L


def default_settings(**kw) -> CSettings:
    s = CSettings()
    load().osc_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def c_spec(spec: RobotSpec) -> CRobotSpec:
    r = CRobotSpec()
    r.nv, r.nu, r.nc, r.ns = spec.nv, spec.nu, spec.nc, spec.ns
    for i in range(spec.ns):
        r.w_trans[i], r.w_rot[i] = spec.w_trans[i], spec.w_rot[i]
    r.w_torque, r.w_reg, r.mu, r.fz_max = spec.w_torque, spec.w_reg, spec.mu, spec.fz_max
    for i in range(spec.nu):
        r.u_lb[i], r.u_ub[i] = spec.u_lb[i], spec.u_ub[i]
    return r


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (osc_host_alloc)."""
    L = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = L.osc_host_alloc(max(n, 16), C.byref(p))
    if rc:
        raise OscError(f"osc_host_alloc failed ({rc}): {L.osc_last_error(None).decode()}")
    buf = (C.c_char * max(n, 16)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = (p, buf)
    return arr


_PINNED = {}

_FIELDS = ("M", "C", "J", "bias", "targets", "mask")


class BatchedOSC:
    """N environments of one robot, one CUDA device.

    Mirrors the reference controller's life cycle
    (walter_sr/operational_space_controller.h:110-240):
        initialize_optimization()      -> setup(inputs)
        control_loop body              -> step(inputs)   (returns torques)
        get_torque_command/get_solution-> torques() / solution() / dual_solution()
    """

    def __init__(self, spec: RobotSpec, n_envs: int, settings: CSettings | None = None,
                 device: int = 0):
        self.L = load()
        self.spec, self.n_envs, self.device = spec, int(n_envs), device
        self.settings = settings if settings is not None else default_settings()
        self._cspec = c_spec(spec)
        self.h = C.c_void_p()
        rc = self.L.osc_create(C.byref(self._cspec), C.byref(self.settings), self.n_envs, device,
                               C.byref(self.h))
        if rc:
            self.h = None
            raise OscError(f"osc_create failed ({rc}): {self.L.osc_last_error(None).decode()}")

    # -- helpers ---------------------------------------------------------
    def _check(self, rc, what):
        if rc:
            raise OscError(f"{what} failed ({rc}): {self.L.osc_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.osc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shapes(self):
        sp, N = self.spec, self.n_envs
        return dict(M=(N, sp.nv, sp.nv), C=(N, sp.nv), J=(N, sp.s, sp.nv), bias=(N, sp.s),
                    targets=(N, sp.ns, 6), mask=(N, sp.nc))

    def _ptrs(self, inputs):
        keep, ptrs = [], []
        shapes = self._shapes()
        for k in _FIELDS:
            a = inputs.get(k) if inputs is not None else None
            if a is None:
                ptrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shapes[k]:
                raise ValueError(f"{k}: expected shape {shapes[k]}, got {a.shape}")
            keep.append(a)
            ptrs.append(a.ctypes.data)
        return keep, ptrs

    # -- API -------------------------------------------------------------
    def upload(self, inputs, stream=None):
        keep, ptrs = self._ptrs(inputs)
        self._check(self.L.osc_upload(self.h, *ptrs, stream), "osc_upload")
        self._check(self.L.osc_sync(self.h, stream), "osc_sync")  # host arrays may go away

    def setup(self, inputs=None, stream=None):
        if inputs is not None:
            self.upload(inputs, stream)
        self._check(self.L.osc_setup(self.h, stream), "osc_setup")

    def step_device(self, stream=None):
        """Control step on the inputs already resident in HBM (no host traffic)."""
        self._check(self.L.osc_step(self.h, stream), "osc_step")

    def step(self, inputs, stream=None) -> np.ndarray:
        """Host buffers in, torques out (upload + step + download + sync)."""
        keep, ptrs = self._ptrs(inputs)
        tq = np.empty((self.n_envs, self.spec.nu))
        self._check(self.L.osc_step_host(self.h, *ptrs, tq.ctypes.data, stream), "osc_step_host")
        return tq

    def step_host_into(self, ptrs, torque_out: np.ndarray, stream=None):
        """Hot-loop variant of step(): pre-resolved input pointers, caller-owned output."""
        self._check(self.L.osc_step_host(self.h, *ptrs, torque_out.ctypes.data, stream),
                    "osc_step_host")

    def step_host_j32_into(self, ptrs, torque_out: np.ndarray, stream=None):
        """step_host_into with the task Jacobian in FP32 (ptrs[2] points at float32 data of the
        same [n_envs, s, nv] shape): opt-in FP32 transport of J, widened on the device."""
        self._check(self.L.osc_step_host_j32(self.h, *ptrs, torque_out.ctypes.data, stream),
                    "osc_step_host_j32")

    def kinematics(self, model: CKinModel, qpos_dev: int, qvel_dev: int, stream=None):
        """M, C, J, bias of every environment from qpos / qvel (DEVICE pointers) into the
        handle's own input buffers (update_mj_data / update_osc_data on the device)."""
        self._check(self.L.osc_kinematics(self.h, C.byref(model), qpos_dev, qvel_dev, stream),
                    "osc_kinematics")

    def step_condensed(self, stream=None):
        """Control step of the CONDENSED fast mode on the inputs resident in HBM (Cholesky of
        M, QP in (u, z) only; reported separately, never the gated path)."""
        self._check(self.L.osc_step_condensed(self.h, stream), "osc_step_condensed")

    def reset_condensed(self, stream=None):
        self._check(self.L.osc_reset_condensed(self.h, stream), "osc_reset_condensed")

    def reset_warm_start(self, stream=None):
        self._check(self.L.osc_reset_warm_start(self.h, stream), "osc_reset_warm_start")

    def sync(self, stream=None):
        self._check(self.L.osc_sync(self.h, stream), "osc_sync")

    def results(self, stream=None) -> dict:
        sp, N = self.spec, self.n_envs
        out = dict(torque=np.empty((N, sp.nu)), x=np.empty((N, sp.n)), y=np.empty((N, sp.m)),
                   iters=np.empty(N, np.int32), status=np.empty(N, np.int32),
                   pri_res=np.empty(N), dua_res=np.empty(N), rho=np.empty(N))
        self._check(self.L.osc_download(
            self.h, out["torque"].ctypes.data, out["x"].ctypes.data, out["y"].ctypes.data,
            out["iters"].ctypes.data, out["status"].ctypes.data, out["pri_res"].ctypes.data,
            out["dua_res"].ctypes.data, out["rho"].ctypes.data, stream), "osc_download")
        self.sync(stream)
        return out

    def objective(self, stream=None):
        """H's dv block [N, nv, nv] and f's dv part [N, nv] as built by the last setup/step."""
        H = np.empty((self.n_envs, self.spec.nv, self.spec.nv))
        f = np.empty((self.n_envs, self.spec.nv))
        self._check(self.L.osc_download_objective(self.h, H.ctypes.data, f.ctypes.data, stream),
                    "osc_download_objective")
        self.sync(stream)
        return H, f

    def torques(self, stream=None) -> np.ndarray:
        tq = np.empty((self.n_envs, self.spec.nu))
        self._check(self.L.osc_download(self.h, tq.ctypes.data, None, None, None, None, None,
                                        None, None, stream), "osc_download")
        self.sync(stream)
        return tq

    def device_buffers(self) -> CDeviceBuffers:
        b = CDeviceBuffers()
        self._check(self.L.osc_get_device_buffers(self.h, C.byref(b)), "osc_get_device_buffers")
        return b

    def bind_device_inputs(self, M=None, C_=None, J=None, bias=None, targets=None, mask=None):
        """Device pointers (ints) of caller-owned HBM buffers to read inputs from."""
        self._check(self.L.osc_bind_device_inputs(self.h, M, C_, J, bias, targets, mask),
                    "osc_bind_device_inputs")

    def targets_pd(self, site_state: dict, kp_lin, kd_lin, kp_ang, kd_ang, stream=None):
        """Task-space PD targets of every (environment, site) computed on the device into the
        handle's `targets` input (examples/standing.cc:146-155).  site_state maps
        pos/quat/vel/angvel/pos_des/quat_des[/vel_des/angvel_des] to DEVICE pointers (ints)."""
        st = CSiteState(**{k: site_state.get(k) for k, _ in CSiteState._fields_})
        g = [np.ascontiguousarray(a, dtype=np.float64) for a in (kp_lin, kd_lin, kp_ang, kd_ang)]
        for a in g:
            if a.shape != (self.spec.ns,):
                raise ValueError(f"gains: expected shape ({self.spec.ns},), got {a.shape}")
        dp = C.POINTER(C.c_double)
        self._check(self.L.osc_targets_pd(self.h, C.byref(st), *[a.ctypes.data_as(dp) for a in g],
                                          stream), "osc_targets_pd")

    def targets_walter_tumbling(self, shin_angle, shin_angle_prev, shin_angle0, thigh_z,
                                thigh_z_prev, thigh_z0, time: float, dt: float, gains=None,
                                stream=None):
        """The Walter tumbling driver's target laws (walter_sr_true_tumbling_mjjoint.cc:695-1019)
        from DEVICE arrays [n_envs, 4] into the handle's `targets` input."""
        self._check(self.L.osc_targets_walter_tumbling(
            self.h, C.byref(gains) if gains is not None else None, shin_angle, shin_angle_prev,
            shin_angle0, thigh_z, thigh_z_prev, thigh_z0, float(time), float(dt), stream),
            "osc_targets_walter_tumbling")

    def contact_mask_from_contacts(self, geom_pairs_dev: int, ncon_dev: int, max_con: int,
                                   contact_geom_ids, site_of_geom=None, stream=None):
        """Contact mask from MuJoCo contact geom pairs resident on the device
        (examples/walter_sr_true_tumbling_mjjoint.cc:523-558) into the handle's `mask` input."""
        ip = C.POINTER(C.c_int)
        ids = np.ascontiguousarray(contact_geom_ids, dtype=np.int32)
        if ids.shape != (self.spec.nc,):
            raise ValueError(f"contact_geom_ids: expected shape ({self.spec.nc},)")
        sog = None
        if site_of_geom is not None:
            sog = np.ascontiguousarray(site_of_geom, dtype=np.int32)
            if sog.shape != ids.shape:
                raise ValueError("site_of_geom: same shape as contact_geom_ids")
        self._check(self.L.osc_contact_mask_from_contacts(
            self.h, geom_pairs_dev, ncon_dev, int(max_con), ids.ctypes.data_as(ip),
            sog.ctypes.data_as(ip) if sog is not None else None, stream),
            "osc_contact_mask_from_contacts")

    # -- multi-GPU gather by peer stores (osc_gather_*) -----------------------
    def gather_create(self, rank: int, world: int) -> bytes:
        """Allocate this rank's gathered slab; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)()
        self._check(self.L.osc_gather_create(self.h, rank, world, buf), "osc_gather_create")
        self._g_world = world
        return bytes(buf)

    def gather_attach(self, ipc_handles: bytes | None = None, peer_slabs=None):
        """Map the peers' slabs: `ipc_handles` = world x 64 bytes in rank order (other
        processes) or `peer_slabs` = device pointers of handles in this process."""
        arr = None
        if peer_slabs is not None:
            arr = (C.c_void_p * len(peer_slabs))(*[C.c_void_p(int(p) if p else 0) for p in peer_slabs])
        raw = None
        if ipc_handles is not None:
            raw = (C.c_ubyte * len(ipc_handles)).from_buffer_copy(ipc_handles)
        self._check(self.L.osc_gather_attach(self.h, raw, arr), "osc_gather_attach")

    def gather_torques(self, stream=None):
        """Push this rank's torques + statistics into every rank's gathered slab."""
        self._check(self.L.osc_gather_torques(self.h, stream), "osc_gather_torques")

    def gather_buffers(self):
        """(torque_all, stats_all) device pointers of this rank's gathered copies."""
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.L.osc_gather_buffers(self.h, C.byref(a), C.byref(b)), "osc_gather_buffers")
        return a.value, b.value

    def host_traffic(self):
        """(h2d_bytes, d2h_bytes) of the last step() / step_host_into()."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        self._check(self.L.osc_host_traffic(self.h, C.byref(a), C.byref(b)), "osc_host_traffic")
        return a.value, b.value

    def set_fused_build(self, on: bool = True):
        """osc_step as two kernels (objective build inside the equilibration kernel; default)
        or three (separate build_qp_kernel).  Same results."""
        self._check(self.L.osc_set_fused_build(self.h, int(on)), "osc_set_fused_build")

    def enable_timing(self, on: bool = True):
        self._check(self.L.osc_timing_enable(self.h, int(on)), "osc_timing_enable")

    def read_timing(self) -> CKernelTimes:
        """Average ms per step of the build and solve kernels (CUDA events on the launch
        stream) since the last read."""
        t = CKernelTimes()
        self._check(self.L.osc_timing_read(self.h, C.byref(t)), "osc_timing_read")
        return t

    def reinit_count(self, stream=None) -> int:
        """Environment-steps that took the sparsity-change (re-Init) path so far."""
        c = C.c_int(0)
        self._check(self.L.osc_reinit_count(self.h, C.byref(c), stream), "osc_reinit_count")
        return c.value

    @property
    def kernel_launches(self) -> int:
        return int(self.L.osc_kernel_launches(self.h))


def measure_dfma_tflops(device: int = 0) -> float:
    v = C.c_double()
    rc = load().osc_measure_dfma_tflops(device, C.byref(v))
    if rc:
        raise OscError(f"osc_measure_dfma_tflops failed ({rc})")
    return v.value
