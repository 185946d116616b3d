"""oracle/osc_condensed.py -- TEST INFRASTRUCTURE (CPU): the CONDENSED fast mode of the
controller, restated with numpy + the OSQP restatement's generic QP entry (osc_oracle.solve_qp).

The reference never forms it: it keeps  M dv + C = B u + Jc z  as equality rows of its QP
(walter_sr/autogen/autogen.py:82-93).  Because dv is unbounded (dv_lb/ub = -/+inf,
walter_sr/operational_space_controller.h:286-287) it can be eliminated exactly,
    dv = G w + d0 ,   w = [u; z] ,   G = M^-1 [B  Jc] ,   d0 = -M^-1 C      (Cholesky of M),
which leaves a QP in (u, z) only
    min 1/2 w'P'w + q'w ,  P' = G' Hd G + R ,  q' = G'(Hd d0 + fd) ,  R = diag(hu.., hz..)
    s.t.  F z <= 0 ,  u_lb <= u <= u_ub ,  z_lb o mask <= z <= z_ub o mask
with the same unique optimum but a different ADMM iterate sequence ("Reconciling
north_star"): a separately reported mode, never the gated path.  Control-step protocol of the
mode (ours to define; it mirrors the reference's re-Init branch :571-584): every step is
osqp_setup on the new condensed data with rho carried over from the previous step, then
osqp_warm_start(previous w, previous dual), then osqp_solve.
"""
from __future__ import annotations

import numpy as np

import osc_oracle as orc

INF = 1e30


def condense(spec, M, C, J, bias, targets, mask):
    """One environment: returns P (n',n'), q, A (m',n'), l, u, G (nv,n'), d0."""
    nv, nu, nc = spec.nv, spec.nu, spec.nc
    nz = 3 * nc
    H, f, A, l, u = orc.build_qp(spec, M, C, J, bias, targets, mask)
    Hd, fd = H[:nv, :nv], f[:nv]
    Aeq = A[:nv]
    Bc = -Aeq[:, nv:]                       # [B | Jc]  (Aeq = [M, -B, -Jc])
    L = np.linalg.cholesky(M)
    G = np.linalg.solve(L.T, np.linalg.solve(L, Bc))
    d0 = -np.linalg.solve(L.T, np.linalg.solve(L, C))
    R = np.diag(np.diag(H)[nv:])
    P = G.T @ Hd @ G + R
    P = 0.5 * (P + P.T)
    q = G.T @ (Hd @ d0 + fd)
    rows = list(range(nv, nv + 4 * nc)) + list(range(nv + 4 * nc + nv, spec.m))
    Ac = A[rows][:, nv:]
    return P, q, Ac, l[rows], u[rows], G, d0


class CondensedOracle:
    """N environments, the condensed mode's control-step protocol (see module docstring)."""

    def __init__(self, spec, n_envs, settings=None):
        self.spec, self.n = spec, n_envs
        self.settings = settings if settings is not None else orc.default_settings()
        self.reset()

    def reset(self):
        self.w = [None] * self.n
        self.y = [None] * self.n
        self.rho = [float(self.settings.rho)] * self.n

    def step(self, inp):
        sp, N = self.spec, self.n
        npr = sp.nu + 3 * sp.nc
        out = dict(torque=np.zeros((N, sp.nu)), w=np.zeros((N, npr)), x=np.zeros((N, sp.n)),
                   y=np.zeros((N, 4 * sp.nc + npr)), iters=np.zeros(N, np.int32),
                   status=np.zeros(N, np.int32), rho=np.zeros(N), pri_res=np.zeros(N),
                   dua_res=np.zeros(N))
        for e in range(N):
            P, q, A, l, u, G, d0 = condense(sp, *[inp[k][e] for k in
                                                  ("M", "C", "J", "bias", "targets", "mask")])
            s = orc.Settings()
            for fld, _ in s._fields_:
                setattr(s, fld, getattr(self.settings, fld))
            s.rho = self.rho[e]
            warm = (self.w[e], self.y[e]) if (self.w[e] is not None and s.warm_start) else None
            r = orc.solve_qp(P, q, A, l, u, s, warm=warm)
            ok = np.isfinite(r["x"]).all()
            self.w[e], self.y[e] = (r["x"], r["y"]) if ok else (None, None)
            self.rho[e] = r["rho"]
            out["w"][e], out["y"][e] = r["x"], r["y"]
            out["torque"][e] = r["x"][:sp.nu]
            out["x"][e] = np.concatenate([G @ r["x"] + d0, r["x"]])
            out["iters"][e], out["status"][e], out["rho"][e] = r["iter"], r["status"], r["rho"]
            out["pri_res"][e], out["dua_res"][e] = r["pri_res"], r["dua_res"]
        return out
