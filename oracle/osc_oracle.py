"""ctypes binding of the CPU oracle (oracle/libosc_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED: see oracle/osqp_restated.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_MAX_SITES = 32
ORC_MAX_NU = 16

STATUS_SOLVED = 1
STATUS_SOLVED_INACCURATE = 2
STATUS_MAX_ITER = -2


class Settings(C.Structure):
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("scaling", C.c_int), ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
        ("max_iter", C.c_int), ("check_termination", C.c_int), ("warm_start", C.c_int),
        ("scaled_termination", C.c_int), ("linsys", C.c_int),
    ]


class Robot(C.Structure):
    _fields_ = [
        ("nv", C.c_int), ("nu", C.c_int), ("nc", C.c_int), ("ns", C.c_int),
        ("w_trans", C.c_double * ORC_MAX_SITES), ("w_rot", C.c_double * ORC_MAX_SITES),
        ("w_torque", C.c_double), ("w_reg", C.c_double), ("mu", C.c_double),
        ("u_lb", C.c_double * ORC_MAX_NU), ("u_ub", C.c_double * ORC_MAX_NU),
        ("fz_max", C.c_double),
    ]


class Info(C.Structure):
    _fields_ = [("iter", C.c_int), ("status", C.c_int), ("pri_res", C.c_double),
                ("dua_res", C.c_double), ("rho_estimate", C.c_double),
                ("rho_updates", C.c_int), ("decision_margin", C.c_double)]


def build(native: bool = False) -> str:
    """Compile the oracle (make).  native=True builds -march=native into a
    separate file for CPU-baseline timing on the machine that runs it."""
    target = "libosc_oracle_native.so" if native else "libosc_oracle.so"
    path = os.path.join(_HERE, target)
    srcs = [os.path.join(_HERE, f) for f in ("osqp_restated.c", "osc_oracle.c",
                                             "osqp_restated.h", "osc_oracle.h")]
    if os.path.exists(path) and all(os.path.getmtime(path) >= os.path.getmtime(s) for s in srcs):
        return path
    arch = "-march=native" if native else "-march=x86-64-v3"
    cmd = ["gcc", "-O3", arch, "-fPIC", "-pthread", "-std=c11", "-ffp-contract=off", "-shared",
           "-o", path, srcs[0], srcs[1], "-lm", "-lpthread"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return path


def lib(native: bool = False):
    global _LIB
    if _LIB is not None and not native:
        return _LIB
    path = os.path.join(_HERE, "libosc_oracle_native.so" if native else "libosc_oracle.so")
    if not os.path.exists(path):
        path = build(native)
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    L.orc_default_settings.argtypes = [C.POINTER(Settings)]
    L.orc_build_qp.argtypes = [C.POINTER(Robot)] + [dp] * 11
    L.orc_batch_create.restype = C.c_void_p
    L.orc_batch_create.argtypes = [C.POINTER(Robot), C.POINTER(Settings), C.c_int]
    L.orc_batch_destroy.argtypes = [C.c_void_p]
    L.orc_batch_setup.argtypes = [C.c_void_p] + [dp] * 6 + [C.c_int]
    L.orc_batch_step.argtypes = ([C.c_void_p] + [dp] * 6 + [C.c_int] + [dp] * 3 + [ip, ip]
                                 + [dp] * 3 + [ip, dp])
    L.orc_batch_scaled_state.argtypes = [C.c_void_p, C.c_int] + [dp] * 6
    L.orc_max_threads.restype = C.c_int
    if not native:
        _LIB = L
    return L


def default_settings(**kw) -> Settings:
    s = Settings()
    lib().orc_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def robot_from_spec(spec) -> Robot:
    r = Robot()
    r.nv, r.nu, r.nc, r.ns = spec.nv, spec.nu, spec.nc, spec.ns
    for i in range(spec.ns):
        r.w_trans[i] = spec.w_trans[i]
        r.w_rot[i] = spec.w_rot[i]
    r.w_torque, r.w_reg, r.mu = spec.w_torque, spec.w_reg, spec.mu
    for i in range(spec.nu):
        r.u_lb[i] = spec.u_lb[i]
        r.u_ub[i] = spec.u_ub[i]
    r.fz_max = spec.fz_max
    return r


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_qp(spec, M, Cv, J, bias, targets, mask):
    """Closed-form QP of ONE environment: returns H (n,n), f, A (m,n), l, u."""
    n, m = spec.n, spec.m
    H = np.zeros(n * n); f = np.zeros(n); A = np.zeros(m * n); l = np.zeros(m); u = np.zeros(m)
    r = robot_from_spec(spec)
    args = [_c(M), _c(Cv), _c(J), _c(bias), _c(targets), _c(mask)]
    lib().orc_build_qp(C.byref(r), *[_p(a) for a in args], _p(H), _p(f), _p(A), _p(l), _p(u))
    return H.reshape(n, n).T.copy(), f, A.reshape(n, m).T.copy(), l, u


class OracleBatch:
    """N independent reference-style controllers (one OSQP workspace each)."""

    def __init__(self, spec, n_envs: int, settings: Settings | None = None, native: bool = False):
        self.spec, self.n_envs = spec, n_envs
        self.L = lib(native)
        self.settings = settings if settings is not None else default_settings()
        self.robot = robot_from_spec(spec)
        self.h = self.L.orc_batch_create(C.byref(self.robot), C.byref(self.settings), n_envs)
        self.max_threads = self.L.orc_max_threads()

    def close(self):
        if self.h:
            self.L.orc_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ins(self, inp):
        arrs = [_c(inp[k]) for k in ("M", "C", "J", "bias", "targets", "mask")]
        assert arrs[0].shape[0] == self.n_envs
        return arrs

    def setup(self, inp, n_threads: int = 0) -> int:
        arrs = self._ins(inp)
        return self.L.orc_batch_setup(self.h, *[_p(a) for a in arrs], n_threads)

    def step(self, inp, n_threads: int = 0) -> dict:
        sp, N = self.spec, self.n_envs
        arrs = self._ins(inp)
        out = dict(torque=np.zeros((N, sp.nu)), x=np.zeros((N, sp.n)), y=np.zeros((N, sp.m)),
                   iters=np.zeros(N, np.int32), status=np.zeros(N, np.int32),
                   pri_res=np.zeros(N), dua_res=np.zeros(N), rho=np.zeros(N),
                   rho_updates=np.zeros(N, np.int32), margin=np.zeros(N))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        out["reinits"] = self.L.orc_batch_step(
            self.h, *[_p(a) for a in arrs], n_threads, _p(out["torque"]), _p(out["x"]),
            _p(out["y"]), ip(out["iters"]), ip(out["status"]), _p(out["pri_res"]),
            _p(out["dua_res"]), _p(out["rho"]), ip(out["rho_updates"]), _p(out["margin"]))
        return out

    def scaled_state(self, e: int) -> dict:
        sp = self.spec
        x = np.zeros(sp.n); z = np.zeros(sp.m); y = np.zeros(sp.m)
        D = np.zeros(sp.n); E = np.zeros(sp.m); c = np.zeros(1)
        self.L.orc_batch_scaled_state(self.h, e, _p(x), _p(z), _p(y), _p(D), _p(E), _p(c))
        return dict(x=x, z=z, y=y, D=D, E=E, c=float(c[0]))


# ---------------------------------------------------------------------------
# generic QP interface of the OSQP restatement (used to pin it on known answers)
# ---------------------------------------------------------------------------
class _Csc(C.Structure):
    _fields_ = [("n_rows", C.c_int), ("n_cols", C.c_int), ("nnz", C.c_int),
                ("p", C.POINTER(C.c_int)), ("i", C.POINTER(C.c_int)), ("x", C.POINTER(C.c_double))]


def solve_qp(P, q, A, l, u, settings: Settings | None = None, warm=None, n_solves: int = 1):
    """min 1/2 x'Px + q'x  s.t. l <= Ax <= u  through orc_setup/orc_solve."""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.orc_csc_from_dense.restype = C.POINTER(_Csc)
    L.orc_csc_from_dense.argtypes = [dp, C.c_int, C.c_int, C.c_int]
    L.orc_csc_free.argtypes = [C.POINTER(_Csc)]
    L.orc_setup.restype = C.c_void_p
    L.orc_setup.argtypes = [C.POINTER(_Csc), dp, C.POINTER(_Csc), dp, dp, C.POINTER(Settings)]
    L.orc_solve.argtypes = [C.c_void_p]
    L.orc_cleanup.argtypes = [C.c_void_p]
    L.orc_warm_start.argtypes = [C.c_void_p, dp, dp]
    L.orc_get_info.restype = C.POINTER(Info)
    L.orc_get_info.argtypes = [C.c_void_p]
    L.orc_solution_x.restype = dp
    L.orc_solution_x.argtypes = [C.c_void_p]
    L.orc_solution_y.restype = dp
    L.orc_solution_y.argtypes = [C.c_void_p]
    L.orc_get_rho.restype = C.c_double
    L.orc_get_rho.argtypes = [C.c_void_p]
    P = np.asarray(P, float); A = np.asarray(A, float)
    q = _c(q); l = _c(l); u = _c(u)
    n, m = P.shape[0], A.shape[0]
    s = settings if settings is not None else default_settings()
    Pc = L.orc_csc_from_dense(_p(np.ascontiguousarray(P.T).ravel()), n, n, 1)
    Ac = L.orc_csc_from_dense(_p(np.ascontiguousarray(A.T).ravel()), m, n, 0)
    w = L.orc_setup(Pc, _p(q), Ac, _p(l), _p(u), C.byref(s))
    L.orc_csc_free(Pc); L.orc_csc_free(Ac)
    if not w:
        raise RuntimeError("orc_setup failed")
    if warm is not None:
        L.orc_warm_start(w, _p(_c(warm[0])), _p(_c(warm[1])))
    for _ in range(n_solves):
        L.orc_solve(w)
    info = L.orc_get_info(w).contents
    out = dict(x=np.array([L.orc_solution_x(w)[i] for i in range(n)]),
               y=np.array([L.orc_solution_y(w)[i] for i in range(m)]),
               iter=info.iter, status=info.status, pri_res=info.pri_res, dua_res=info.dua_res,
               rho=L.orc_get_rho(w), rho_updates=info.rho_updates)
    L.orc_cleanup(w)
    return out
