"""CPU tests that PIN THE ORACLE (the reference ships no tests for this path).

1. OSQP restatement vs the solver log printed in OSQP's own documentation.
2. Closed-form H, f, Aeq, beq, Aineq, bineq vs a literal transcription of the reference's
   autogen.py differentiated numerically.
3. Solver-independent certificates: KKT conditions, active-set re-solve, scipy SLSQP.
4. Analytic case: all contacts masked out.
5. Drift guard against tests/golden/oracle_cases.npz.
"""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("M", "C", "J", "bias", "targets", "mask")


def _env(inp, e):
    return [inp[k][e] for k in FIELDS]


# --------------------------------------------------------------------------- 1
def test_osqp_restatement_reproduces_published_demo_log(oracle):
    g = json.load(open(os.path.join(HERE, "golden", "osqp_demo.json")))
    for linsys in (0, 1):
        kw = dict(linsys=linsys, **g["settings"])
        r1 = oracle.solve_qp(g["P"], g["q"], g["A"], g["l"], g["u"],
                             oracle.default_settings(max_iter=1, **kw))
        x = r1["x"]
        obj = 0.5 * x @ np.array(g["P"]) @ x + np.array(g["q"]) @ x
        assert f"{obj:.4e}" == g["log_iter_1"]["objective"]
        assert f"{r1['pri_res']:.2e}" == g["log_iter_1"]["pri_res"]
        assert f"{r1['dua_res']:.2e}" == g["log_iter_1"]["dua_res"]
        r = oracle.solve_qp(g["P"], g["q"], g["A"], g["l"], g["u"], oracle.default_settings(**kw))
        x = r["x"]
        obj = 0.5 * x @ np.array(g["P"]) @ x + np.array(g["q"]) @ x
        assert r["iter"] == g["number_of_iterations"] and r["status"] == oracle.STATUS_SOLVED
        assert f"{obj:.4e}" == g["log_iter_50"]["objective"]
        assert f"{r['pri_res']:.2e}" == g["log_iter_50"]["pri_res"]
        assert f"{r['dua_res']:.2e}" == g["log_iter_50"]["dua_res"]
        assert f"{r['rho']:.2e}" == g["log_iter_50"]["rho"]
        np.testing.assert_allclose(r["x"], g["x"], atol=1e-6)
        np.testing.assert_allclose(r["y"], g["y"], atol=1e-6)


def test_osqp_restatement_small_known_answers(oracle):
    # box-constrained scalar: min 1/2 x^2 - 3x, 0<=x<=1  -> x=1, y=2
    r = oracle.solve_qp([[1.0]], [-3.0], [[1.0]], [0.0], [1.0],
                        oracle.default_settings(eps_abs=1e-9, eps_rel=1e-9))
    np.testing.assert_allclose(r["x"], [1.0], atol=1e-7)
    np.testing.assert_allclose(r["y"], [2.0], atol=1e-6)
    # equality constrained: min x1^2+x2^2 s.t. x1+x2=1 -> (0.5,0.5), y=-1
    r = oracle.solve_qp(2 * np.eye(2), [0.0, 0.0], [[1.0, 1.0]], [1.0], [1.0],
                        oracle.default_settings(eps_abs=1e-9, eps_rel=1e-9))
    np.testing.assert_allclose(r["x"], [0.5, 0.5], atol=1e-7)
    np.testing.assert_allclose(r["y"], [-1.0], atol=1e-6)
    # infinite bounds are clipped to 1e30 and typed "loose"
    r = oracle.solve_qp(np.eye(2), [1.0, -1.0], np.eye(2), [-1e40, -1e40], [1e40, 0.25],
                        oracle.default_settings(eps_abs=1e-9, eps_rel=1e-9))
    np.testing.assert_allclose(r["x"], [-1.0, 0.25], atol=1e-7)
    # x >= 1 and x <= 0: primal infeasible (status -3), NaN solution (store_solution)
    r = oracle.solve_qp([[1.0]], [0.0], [[1.0], [1.0]], [1.0, -1e40], [1e40, 0.0])
    assert r["status"] == -3 and np.isnan(r["x"]).all() and np.isnan(r["y"]).all()
    # min -x, x >= 0, no curvature: unbounded below = dual infeasible (status -4)
    r = oracle.solve_qp([[0.0]], [-1.0], [[1.0]], [0.0], [1e40])
    assert r["status"] == -4 and np.isnan(r["x"]).all()


# --------------------------------------------------------------------------- 2
# literal transcription of <robot>/autogen/autogen.py (symbolic -> numeric)
SITE_KEYS = {  # walter_sr/autogen/autogen.py:167-168 ; unitree_go2/autogen/autogen.py:160-161
    17: ["torso", "tls", "trs", "hls", "hrs", "tlh", "trh", "hlh", "hrh",
         "tlf", "tlr", "trf", "trr", "hlf", "hlr", "hrf", "hrr"],
    5: ["base", "fr", "fl", "hr", "hl"],
}


def _literal(spec):
    nv, nu, nz, ns = spec.nv, spec.nu, spec.nz, spec.ns
    dv_idx, u_idx, z_idx = nv, nv + nu, nv + nu + nz
    B = np.vstack([np.zeros((nv - nu, nu)), np.eye(nu)])  # autogen.py:54-60
    keys = SITE_KEYS[ns]
    weights = {}
    for i, k in enumerate(keys):
        weights[f"{k}_translational_tracking"] = spec.w_trans[i]
        weights[f"{k}_rotational_tracking"] = spec.w_rot[i]
    weights["torque"], weights["regularization"] = spec.w_torque, spec.w_reg

    def equality(q, M, C, Jc):  # autogen.py:62-93
        dv, u, z = q[:dv_idx], q[dv_idx:u_idx], q[u_idx:z_idx]
        return M @ dv + C - B @ u - Jc @ z

    def inequality(q):  # autogen.py:95-133
        z = q[u_idx:z_idx]
        out = []
        for x in np.split(z, spec.nc):
            out += [x[0] + x[1] - spec.mu * x[2], -x[0] + x[1] - spec.mu * x[2],
                    x[0] - x[1] - spec.mu * x[2], -x[0] - x[1] - spec.mu * x[2]]
        return np.array(out)

    def objective(q, desired, J, bias):  # autogen.py:135-345
        dv, u = q[:dv_idx], q[dv_idx:u_idx]
        ddx = J @ dv + bias
        ddx_p, ddx_r = np.split(ddx, 2)
        ddx_p, ddx_r = np.split(ddx_p, ns), np.split(ddx_r, ns)
        des_p, des_r = desired[:, :3], desired[:, 3:]
        terms = {}
        for i, k in enumerate(keys):
            terms[f"{k}_translational_tracking"] = np.sum((ddx_p[i] - des_p[i]) ** 2)
            terms[f"{k}_rotational_tracking"] = np.sum((ddx_r[i] - des_r[i]) ** 2)
        terms["torque"] = np.sum(u ** 2)
        terms["regularization"] = np.sum(q ** 2)
        return sum(v * weights[k] for k, v in terms.items())

    return equality, inequality, objective


@pytest.mark.parametrize("preset,config", [("walter_sr_true_tumbling_mjjoint", "tumbling"),
                                           ("unitree_go2", "go2_standing"),
                                           ("walter_sr_wheels", "stairs")])
def test_closed_form_qp_matches_literal_autogen(oracle, preset, config):
    import osc_b200 as ob
    spec = ob.load_preset(preset)
    inp = ob.synth.make_inputs(spec, 4, config)
    equality, inequality, objective = _literal(spec)
    n = spec.n
    I = np.eye(n)
    for e in range(2):
        M, C, J, bias, targets, mask = _env(inp, e)
        H, f, A, l, u = oracle.build_qp(spec, M, C, J, bias, targets, mask)
        Jc = J[3 * spec.ns - spec.nz:3 * spec.ns].T  # reference :497-503
        z0 = np.zeros(n)
        # the objective is exactly quadratic: differences at unit steps are exact
        f0 = objective(z0, targets, J, bias)
        fi = np.array([objective(I[i], targets, J, bias) for i in range(n)])
        Hn = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                Hn[i, j] = Hn[j, i] = objective(I[i] + I[j], targets, J, bias) - fi[i] - fi[j] + f0
        gn = fi - f0 - 0.5 * np.diag(Hn)
        scale = max(1.0, np.abs(H).max())
        np.testing.assert_allclose(H, Hn, atol=2e-9 * scale)
        np.testing.assert_allclose(f, gn, atol=2e-9 * max(1.0, np.abs(f).max(), scale))
        # Aeq = d eq / dq (affine -> exact), beq = -eq(0)
        eq0 = equality(z0, M, C, Jc)
        Aeq = np.stack([equality(I[i], M, C, Jc) - eq0 for i in range(n)], axis=1)
        np.testing.assert_allclose(A[:spec.nv], Aeq, atol=1e-12 * max(1, np.abs(Aeq).max()))
        np.testing.assert_array_equal(l[:spec.nv], -eq0)
        np.testing.assert_array_equal(u[:spec.nv], -eq0)
        in0 = inequality(z0)
        Ain = np.stack([inequality(I[i]) - in0 for i in range(n)], axis=1)
        np.testing.assert_array_equal(A[spec.nv:spec.nv + 4 * spec.nc], Ain)
        np.testing.assert_array_equal(u[spec.nv:spec.nv + 4 * spec.nc], -in0)
        np.testing.assert_array_equal(A[spec.nv + 4 * spec.nc:], np.eye(n))  # Abox (:284-285)


def test_bounds_and_contact_indexing(oracle):
    """contact c <-> z[3c:3c+3] <-> friction rows nv+4c.. <-> box rows <-> mask[c] (:546-555)."""
    import osc_b200 as ob
    spec = ob.load_preset("walter_sr")
    inp = ob.synth.make_inputs(spec, 1, "tumbling")
    M, C, J, bias, targets, _ = _env(inp, 0)
    nv, nu, nc, n = spec.nv, spec.nu, spec.nc, spec.n
    for c in range(nc):
        mask = np.ones(nc)
        mask[c] = 0.0
        _, _, A, l, u = oracle.build_qp(spec, M, C, J, bias, targets, mask)
        rb = nv + 4 * nc + nv + nu
        for cc in range(nc):
            lo, hi = l[rb + 3 * cc:rb + 3 * cc + 3], u[rb + 3 * cc:rb + 3 * cc + 3]
            if cc == c:
                assert np.all(lo == 0.0) and np.all(hi == 0.0)
            else:
                np.testing.assert_array_equal(lo, [-1e30, -1e30, 0.0])
                np.testing.assert_array_equal(hi, [1e30, 1e30, 1e4])
        fr = A[nv + 4 * c:nv + 4 * c + 4]
        assert np.all(fr[:, :nv + nu + 3 * c] == 0) and np.all(fr[:, nv + nu + 3 * c + 3:] == 0)
        np.testing.assert_array_equal(fr[:, nv + nu + 3 * c:nv + nu + 3 * c + 3],
                                      [[1, 1, -spec.mu], [-1, 1, -spec.mu],
                                       [1, -1, -spec.mu], [-1, -1, -spec.mu]])
    np.testing.assert_array_equal(l[nv + 4 * nc + nv:nv + 4 * nc + nv + nu], spec.u_lb)
    spec_g = ob.load_preset("unitree_go2")
    assert spec_g.u_ub[:3] == (23.7, 23.7, 45.3)


# --------------------------------------------------------------------------- 3
def _kkt(H, f, A, l, u, x, y):
    stat = np.abs(H @ x + f + A.T @ y).max()
    Ax = A @ x
    pinf = max((l - Ax).max(), (Ax - u).max(), 0.0)
    # dual sign: y_i > 0 only at the upper bound, < 0 only at the lower bound
    tol = 1e-5 * (1 + np.abs(Ax))
    bad_pos = (y > 1e-6) & (np.abs(Ax - u) > tol)
    bad_neg = (y < -1e-6) & (np.abs(Ax - l) > tol)
    return stat, pinf, int(bad_pos.sum() + bad_neg.sum())


@pytest.mark.parametrize("preset,config", [("walter_sr_true_tumbling_mjjoint", "tumbling"),
                                           ("unitree_go2", "go2_standing"),
                                           ("walter_sr_wheels", "stairs"),
                                           ("walter_sr", "standing")])
def test_tight_solution_satisfies_kkt_and_active_set_resolve(oracle, preset, config):
    import osc_b200 as ob
    spec = ob.load_preset(preset)
    N = 16
    inp = ob.synth.make_inputs(spec, N, config)
    b = oracle.OracleBatch(spec, N, oracle.default_settings(eps_abs=1e-10, eps_rel=1e-10,
                                                            max_iter=400000, linsys=1))
    b.setup(inp)
    o = b.step(inp)
    assert (o["status"] == oracle.STATUS_SOLVED).all()
    for e in range(N):
        H, f, A, l, u = oracle.build_qp(spec, *_env(inp, e))
        x, y = o["x"][e], o["y"][e]
        stat, pinf, bad = _kkt(H, f, A, l, u, x, y)
        sc = 1 + np.abs(f).max()
        assert stat < 1e-6 * sc and pinf < 1e-6 and bad == 0
        # independent re-solve: equality-constrained QP on the identified active set
        Ax = A @ x
        act = (np.abs(y) > 1e-7) | (u - l < 1e-9)
        rhs_b = np.where(np.abs(Ax - l) <= np.abs(Ax - u), l, u)[act]
        Aa = A[act]
        K = np.block([[H, Aa.T], [Aa, np.zeros((Aa.shape[0],) * 2)]])
        sol = np.linalg.lstsq(K, np.concatenate([-f, rhs_b]), rcond=None)[0]
        tq, tq2 = x[spec.nv:spec.nv + spec.nu], sol[spec.nv:spec.nv + spec.nu]
        # the torque/contact-force split is only pinned by the 1e-4 regulariser
        # (cond(H) ~ 5e7), so residuals of 1e-10 leave ~1e-4 relative play
        np.testing.assert_allclose(tq2, tq, rtol=2e-3, atol=5e-4 * max(1.0, np.abs(tq).max()))


def test_scipy_slsqp_agrees_on_objective(oracle):
    import osc_b200 as ob
    from scipy.optimize import minimize
    spec = ob.load_preset("unitree_go2")
    inp = ob.synth.make_inputs(spec, 2, "go2_standing")
    b = oracle.OracleBatch(spec, 2, oracle.default_settings(eps_abs=1e-10, eps_rel=1e-10,
                                                            max_iter=400000, linsys=1))
    b.setup(inp)
    o = b.step(inp)
    for e in range(2):
        H, f, A, l, u = oracle.build_qp(spec, *_env(inp, e))
        x = o["x"][e]
        obj = lambda v: 0.5 * v @ H @ v + f @ v
        eq = u - l < 1e-9
        fin_u = (u < 1e29) & ~eq
        fin_l = (l > -1e29) & ~eq
        cons = [dict(type="eq", fun=lambda v: A[eq] @ v - l[eq], jac=lambda v: A[eq]),
                dict(type="ineq", fun=lambda v: u[fin_u] - A[fin_u] @ v, jac=lambda v: -A[fin_u]),
                dict(type="ineq", fun=lambda v: A[fin_l] @ v - l[fin_l], jac=lambda v: A[fin_l])]
        r = minimize(obj, x + 0.1, jac=lambda v: H @ v + f, constraints=cons, method="SLSQP",
                     options=dict(maxiter=500, ftol=1e-12))
        assert r.success
        assert abs(r.fun - obj(x)) <= 1e-5 * (1 + abs(obj(x)))
        np.testing.assert_allclose(r.x[spec.nv:spec.nv + spec.nu], x[spec.nv:spec.nv + spec.nu],
                                   atol=2e-3)


# --------------------------------------------------------------------------- 4
def test_all_contacts_masked_closed_form(oracle):
    """mask = 0 => z in [0,0]; with no torque bound active the QP is an equality-constrained
    least squares:  min 1/2 x'Hx + f'x  s.t.  M dv - B u = -C."""
    import osc_b200 as ob
    spec = ob.load_preset("walter_sr_true_tumbling_mjjoint")
    N = 8
    inp = ob.synth.make_inputs(spec, N, "standing")
    inp["mask"][:] = 0.0
    b = oracle.OracleBatch(spec, N, oracle.default_settings(eps_abs=1e-10, eps_rel=1e-10,
                                                            max_iter=400000, linsys=1))
    b.setup(inp)
    o = b.step(inp)
    nv, nu = spec.nv, spec.nu
    for e in range(N):
        H, f, A, l, u = oracle.build_qp(spec, *_env(inp, e))
        nn = nv + nu
        K = np.block([[H[:nn, :nn], A[:nv, :nn].T], [A[:nv, :nn], np.zeros((nv, nv))]])
        sol = np.linalg.solve(K, np.concatenate([-f[:nn], l[:nv]]))
        assert np.abs(sol[nv:nn]).max() < 1000.0  # torque bounds inactive
        np.testing.assert_allclose(o["x"][e][nn:], 0.0, atol=1e-8)
        np.testing.assert_allclose(o["x"][e][:nn], sol[:nn], rtol=1e-6, atol=1e-6)


# --------------------------------------------------------------------------- 5
def test_oracle_matches_committed_golden_cases(oracle):
    import osc_b200 as ob
    g = np.load(os.path.join(HERE, "golden", "oracle_cases.npz"))
    cases = sorted({tuple(k.split("|")[:2]) for k in g.files})
    assert len(cases) == 4
    for preset, config in cases:
        spec = ob.load_preset(preset)
        b = oracle.OracleBatch(spec, 32, oracle.default_settings())
        for t in range(3):
            inp = ob.synth.make_inputs(spec, 32, config, step=t)
            h = hashlib.sha256()
            for k in FIELDS:
                h.update(np.ascontiguousarray(inp[k]).tobytes())
            key = f"{preset}|{config}|{t}"
            assert h.hexdigest() == bytes(g[key + "|sha"]).decode(), "synthetic inputs drifted"
            if t == 0:
                assert b.setup(inp) == 0
            o = b.step(inp)
            np.testing.assert_array_equal(o["iters"], g[key + "|iters"])
            np.testing.assert_array_equal(o["status"], g[key + "|status"])
            np.testing.assert_allclose(o["torque"], g[key + "|torque"], rtol=1e-9, atol=1e-9)


def test_two_linear_solvers_agree_and_warm_start_helps(oracle):
    """KKT LDL' (OSQP's form) and the reduced Cholesky form are the same algorithm."""
    import osc_b200 as ob
    spec = ob.load_preset("walter_sr_true_tumbling_mjjoint")
    N = 64
    s0 = ob.synth.make_inputs(spec, N, "tumbling", step=0)
    s1 = ob.synth.make_inputs(spec, N, "tumbling", step=1)
    outs = []
    for linsys in (0, 1):
        b = oracle.OracleBatch(spec, N, oracle.default_settings(linsys=linsys))
        b.setup(s0)
        outs.append((b.step(s0), b.step(s1)))
    for a, c in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a["iters"], c["iters"])
        d = np.abs(a["torque"] - c["torque"])
        assert (d <= 0.25 * (1e-5 + 1e-4 * np.abs(a["torque"]))).all()
    assert outs[0][1]["iters"].mean() < outs[0][0]["iters"].mean()
    assert outs[0][1]["reinits"] == 0


def test_sparsity_change_takes_reinit_path(oracle):
    """reference :565-584: a changed sparsity pattern re-Inits the solver and warm starts it
    from the previous (unscaled) solution."""
    import osc_b200 as ob
    spec = ob.load_preset("unitree_go2")
    N = 4
    s0 = ob.synth.make_inputs(spec, N, "go2_standing", step=0)
    s1 = {k: v.copy() for k, v in ob.synth.make_inputs(spec, N, "go2_standing", step=1).items()}
    s1["M"][:, 0, 1] = s1["M"][:, 1, 0] = 1e-3  # a structural zero becomes non-zero
    b = oracle.OracleBatch(spec, N, oracle.default_settings())
    b.setup(s0)
    b.step(s0)
    o = b.step(s1)
    assert o["reinits"] == N and (o["status"] == oracle.STATUS_SOLVED).all()


@pytest.mark.parametrize("preset,config", [("walter_sr_true_tumbling_mjjoint", "tumbling"),
                                           ("unitree_go2", "go2_standing")])
def test_duals_of_the_unbounded_identity_rows_are_exactly_zero(oracle, preset, config):
    """The identity rows of the dv variables have bounds -/+ OSQP_INFTY (rho = RHO_MIN, no
    projection): z_new = z_relaxed + y / rho, y += rho (z_relaxed - z_new) keeps y == 0
    exactly from a cold start on, in every linear-solver form, over cold and warm steps and in
    the scaled iterates.  The two-lane iteration of the device (osc_core3.cuh iterate_pair)
    relies on it: it carries neither y nor rho of these rows."""
    import osc_b200 as ob
    spec = ob.load_preset(preset)
    n_envs = 64
    rb = spec.nv + 4 * spec.nc  # first identity row
    for linsys in (0, 1):
        b = oracle.OracleBatch(spec, n_envs, oracle.default_settings(linsys=linsys))
        b.setup(ob.synth.make_inputs(spec, n_envs, config, step=0))
        for t in range(3):
            o = b.step(ob.synth.make_inputs(spec, n_envs, config, step=t))
            ok = o["status"] == 1
            assert ok.any()
            assert not np.any(o["y"][ok, rb:rb + spec.nv]), (linsys, t)
            assert not np.any(b.scaled_state(0)["y"][rb:rb + spec.nv])
