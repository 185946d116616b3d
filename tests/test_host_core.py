"""CPU: the product's per-environment algorithm (csrc/osc_core3.cuh), run by tests/host_core
on an emulated 32-lane warp (csrc/osc_warp.cuh), against the oracle.  This checks the *device
code's* arithmetic (structured KKT elimination, scaling, rho rules, termination, the lane
mappings for both robot shapes) without a GPU; the GPU tests check the same code as the
kernels actually run it."""
import ctypes as C

import numpy as np
import pytest

FIELDS = ("M", "C", "J", "bias", "targets", "mask")
ATOL, RTOL = 1e-5, 1e-4


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def run_host_core(L, spec, settings, steps, n_envs):
    from osc_b200 import capi
    cs = capi.c_spec(spec)
    SS = L.osc_core_host_state_size(C.byref(cs))
    assert SS >= spec.n + 2 * spec.m + spec.nv + 2
    outs = [dict(torque=np.zeros((n_envs, spec.nu)), iters=np.zeros(n_envs, np.int32),
                 status=np.zeros(n_envs, np.int32), rho=np.zeros(n_envs),
                 reinit=np.zeros(n_envs, np.int32),
                 x=np.zeros((n_envs, spec.n)), y=np.zeros((n_envs, spec.m))) for _ in steps]
    ii = np.zeros(4, np.int32)
    dd = np.zeros(3)
    for e in range(n_envs):
        state = np.zeros(SS)
        a = [np.ascontiguousarray(steps[0][k][e]) for k in FIELDS]
        # osc_setup
        L.osc_core_host_step(C.byref(cs), C.byref(settings), *[_p(v) for v in a], _p(state),
                             None, None, None, None, None, None, None)
        x = np.zeros(spec.n); y = np.zeros(spec.m); tq = np.zeros(spec.nu)
        for t, data in enumerate(steps):
            a = [np.ascontiguousarray(data[k][e]) for k in FIELDS]
            L.osc_core_host_step(C.byref(cs), C.byref(settings), *[_p(v) for v in a], _p(state),
                                 _p(x), _p(y), _p(tq), ii.ctypes.data_as(C.POINTER(C.c_int)),
                                 _p(dd), None, None)
            o = outs[t]
            o["torque"][e], o["x"][e], o["y"][e] = tq, x, y
            o["iters"][e], o["status"][e], o["rho"][e], o["reinit"][e] = ii[0], ii[1], dd[2], ii[3]
    return outs


@pytest.mark.parametrize("preset,config", [("walter_sr_true_tumbling_mjjoint", "tumbling"),
                                           ("unitree_go2", "go2_standing"),
                                           ("walter_sr_wheels", "stairs")])
def test_device_algorithm_matches_oracle_on_host(oracle, host_core, preset, config):
    import osc_b200 as ob
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 48
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(2)]
    b = oracle.OracleBatch(spec, N, oracle.default_settings())
    b.setup(steps[0])
    ref = [b.step(s) for s in steps]
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
    got = run_host_core(host_core, spec, st, steps, N)
    for t in range(2):
        o, g = ref[t], got[t]
        keep = o["margin"] > 1e-7
        np.testing.assert_array_equal(g["iters"][keep], o["iters"][keep])
        np.testing.assert_array_equal(g["status"][keep], o["status"][keep])
        d = np.abs(g["torque"] - o["torque"])[keep]
        tol = (ATOL + RTOL * np.abs(o["torque"]))[keep]
        assert (d <= tol).all(), (preset, t, (d / tol).max())
    assert ref[1]["iters"].mean() < ref[0]["iters"].mean()  # warm start carried over


def test_objective_build_is_bitwise_close_to_oracle(oracle, host_core):
    """H (dv block) and f of the build kernel's per-entry functions vs orc_build_qp."""
    import osc_b200 as ob
    from osc_b200 import capi
    for preset, config in (("walter_sr", "tumbling"), ("unitree_go2", "go2_standing")):
        spec = ob.load_preset(preset)
        inp = ob.synth.make_inputs(spec, 3, config)
        cs = capi.c_spec(spec)
        st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
        for e in range(3):
            a = [np.ascontiguousarray(inp[k][e]) for k in FIELDS]
            H = np.zeros(spec.nv ** 2)
            f = np.zeros(spec.nv)
            host_core.osc_core_host_step(C.byref(cs), C.byref(st), *[_p(v) for v in a], None,
                                         None, None, None, None, None, _p(H), _p(f))
            Ho, fo, *_ = oracle.build_qp(spec, *a)
            np.testing.assert_allclose(H.reshape(spec.nv, spec.nv), Ho[:spec.nv, :spec.nv],
                                       rtol=1e-13, atol=1e-10)
            np.testing.assert_allclose(f, fo[:spec.nv], rtol=1e-13, atol=1e-9)
            assert np.all(fo[spec.nv:] == 0.0)
            np.testing.assert_array_equal(np.diag(Ho)[spec.nv:spec.nv + spec.nu],
                                          2 * (spec.w_reg + spec.w_torque))


@pytest.mark.parametrize("preset,config", [("unitree_go2", "go2_standing"),
                                           ("walter_sr_true_tumbling_mjjoint", "tumbling")])
def test_sparsity_change_branch_matches_oracle(oracle, host_core, preset, config):
    """A structural zero of M becomes non-zero between steps: the reference re-Inits OSQP
    (rho back to settings.rho, scaling with the current cost) and warm starts from the
    previous unscaled solution (:571-584).  Same on the device code."""
    import osc_b200 as ob
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 16
    s0 = ob.synth.make_inputs(spec, N, config, step=0)
    s1 = {k: v.copy() for k, v in ob.synth.make_inputs(spec, N, config, step=1).items()}
    assert (s1["M"][:, 0, 1] == 0).all()
    s1["M"][:, 0, 1] = s1["M"][:, 1, 0] = 1e-3
    s2 = ob.synth.make_inputs(spec, N, config, step=2)  # pattern changes back
    steps = [s0, s1, s2]
    b = oracle.OracleBatch(spec, N, oracle.default_settings())
    b.setup(s0)
    ref = [b.step(s) for s in steps]
    assert [r["reinits"] for r in ref] == [0, N, N]
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
    got = run_host_core(host_core, spec, st, steps, N)
    for t in range(3):
        assert got[t]["reinit"].sum() == ref[t]["reinits"]
        np.testing.assert_array_equal(got[t]["iters"], ref[t]["iters"])
        d = np.abs(got[t]["torque"] - ref[t]["torque"])
        assert (d <= ATOL + RTOL * np.abs(ref[t]["torque"])).all(), (t, d.max())


SETTINGS_VARIANTS = [
    dict(scaling=0),
    dict(scaling=3),
    dict(adaptive_rho=0),
    dict(warm_start=0),
    dict(check_termination=10),
    dict(alpha=1.0, rho=1.0),
    dict(check_termination=0, max_iter=60),
    dict(adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5),
    dict(sigma=1e-4, max_iter=40),
    # events that do not line up: rho adaptation between termination checks, a budget that
    # ends between both (the stretches of the ADMM loop end at whichever comes first)
    dict(adaptive_rho_interval=30, eps_abs=1e-6, eps_rel=1e-6),
    dict(check_termination=7, adaptive_rho_interval=10, max_iter=45, eps_abs=1e-7, eps_rel=1e-7),
    dict(check_termination=25, adaptive_rho_interval=25, max_iter=25),
    dict(max_iter=1),
]


@pytest.mark.parametrize("preset,config", [("walter_sr_wheels", "stairs"),
                                           ("unitree_go2", "go2_standing")])
@pytest.mark.parametrize("kw", SETTINGS_VARIANTS, ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()))
def test_settings_variants_match_oracle(oracle, host_core, preset, config, kw):
    """OsqpSettings other than the defaults (the reference passes the struct through, :110):
    scaling passes, rho adaptation on/off and its interval, warm start off, termination-check
    interval (incl. never), relaxation, step sizes, iteration cap -- device code vs oracle."""
    import osc_b200 as ob
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 12
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(2)]
    b = oracle.OracleBatch(spec, N, oracle.default_settings(**kw))
    b.setup(steps[0])
    ref = [b.step(s) for s in steps]
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
    for k, v in kw.items():
        setattr(st, k, v)
    got = run_host_core(host_core, spec, st, steps, N)
    for t in range(2):
        o, g = ref[t], got[t]
        keep = o["margin"] > 1e-7
        assert keep.mean() > 0.8
        np.testing.assert_array_equal(g["iters"][keep], o["iters"][keep])
        np.testing.assert_array_equal(g["status"][keep], o["status"][keep])
        d = np.abs(g["torque"] - o["torque"])[keep]
        tol = (ATOL + RTOL * np.abs(o["torque"]))[keep]
        assert (d <= tol).all(), (preset, kw, t, (d / tol).max())


@pytest.mark.parametrize("preset,config", [("walter_sr_true_tumbling_mjjoint", "tumbling"),
                                           ("unitree_go2", "go2_standing")])
def test_no_cross_lane_hazard_inside_a_body(host_core, host_core_reversed, preset, config):
    """Race check without a GPU: osc_core3.cuh promises that inside one OSC_LANES body no lane
    reads what another lane writes (on the device the lanes of a body run concurrently, with
    barriers only between bodies).  Running the emulated lanes in the opposite order must then
    give bit-identical results, including on the re-Init path and with rho updates."""
    import osc_b200 as ob
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 10
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(3)]
    steps[1] = {k: v.copy() for k, v in steps[1].items()}
    steps[1]["M"][::2, 0, 1] = steps[1]["M"][::2, 1, 0] = 1e-3   # sparsity change: re-Init path
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 25, 4000, 25, 1)  # rho update every 25
    a = run_host_core(host_core, spec, st, steps, N)
    b = run_host_core(host_core_reversed, spec, st, steps, N)
    for t in range(3):
        for k in ("torque", "x", "y", "iters", "status", "rho", "reinit"):
            assert np.array_equal(a[t][k], b[t][k]), (preset, t, k)
    assert (a[0]["rho"] != 0.1).any()      # rho updates + refactorisations happened
    assert a[1]["reinit"].sum() == N // 2  # and so did the re-Init path


def _infeasible_steps(ob, spec, config, N):
    """Step 1 of the even environments is primal infeasible: a zero row/column in the mass
    matrix of an unactuated dof, no contacts, a large bias force on that dof (0 = -C_0)."""
    steps = [{k: v.copy() for k, v in ob.synth.make_inputs(spec, N, config, step=t).items()}
             for t in range(4)]
    s = steps[1]
    s["M"][::2, 0, :] = 0
    s["M"][::2, :, 0] = 0
    s["mask"][::2] = 0
    s["C"][::2, 0] = 1e4
    return steps


@pytest.mark.parametrize("preset,config", [("walter_sr", "tumbling"), ("unitree_go2", "go2_standing")])
def test_primal_infeasibility_certificate_matches_oracle(oracle, host_core, preset, config):
    """OSQP's primal infeasibility certificate (util.c is_primal_infeasible) and what follows
    it in the reference: status -3, NaN solution / torque (store_solution), then -- because the
    data's sparsity pattern changes back -- a re-Init warm started from that NaN solution
    (:571-584), whose norms ignore NaN (vec_norm_inf), so the environment reports "solved" with
    NaN torques from then on.  Same on the device code, while its neighbours are untouched."""
    import osc_b200 as ob
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 8
    steps = _infeasible_steps(ob, spec, config, N)
    b = oracle.OracleBatch(spec, N, oracle.default_settings())
    b.setup(steps[0])
    ref = [b.step(s) for s in steps]
    assert (ref[1]["status"][::2] == -3).all() and (ref[1]["status"][1::2] == 1).all()
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
    got = run_host_core(host_core, spec, st, steps, N)
    for t in range(4):
        o, g = ref[t], got[t]
        np.testing.assert_array_equal(g["status"], o["status"])
        np.testing.assert_array_equal(g["iters"], o["iters"])
        np.testing.assert_array_equal(np.isnan(g["torque"]), np.isnan(o["torque"]))
        np.testing.assert_array_equal(np.isnan(g["y"]), np.isnan(o["y"]))
        fin = ~np.isnan(o["torque"])
        d = np.abs(g["torque"] - o["torque"])[fin]
        assert (d <= (ATOL + RTOL * np.abs(o["torque"]))[fin]).all(), (preset, t, d.max())
    assert np.isnan(ref[3]["torque"][::2]).all() and not np.isnan(ref[3]["torque"][1::2]).any()


@pytest.mark.parametrize("preset,config", [("walter_sr_wheels", "stairs"), ("unitree_go2", "go2_standing")])
def test_infeasible_step_on_the_update_path_recovers_like_oracle(oracle, host_core, preset, config):
    """Same sparsity pattern throughout (osqp_update_P_A path): an infeasible step leaves NaN
    outputs and zeroed iterates (store_solution -> cold_start), the next feasible step solves
    from cold with the rho the infeasible solve ended with."""
    import osc_b200 as ob
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 6
    steps = [{k: v.copy() for k, v in ob.synth.make_inputs(spec, N, config, step=t).items()}
             for t in range(4)]
    for t, s in enumerate(steps):
        s["M"][:, 0, :] = 0
        s["M"][:, :, 0] = 0
        s["mask"][:] = 0
        s["C"][:, 0] = 0.0
    steps[2]["C"][:3, 0] = 1e4
    b = oracle.OracleBatch(spec, N, oracle.default_settings())
    b.setup(steps[0])
    ref = [b.step(s) for s in steps]
    assert [r["reinits"] for r in ref] == [0, 0, 0, 0]
    assert (ref[2]["status"][:3] == -3).all() and (ref[3]["status"] == 1).all()
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
    got = run_host_core(host_core, spec, st, steps, N)
    for t in range(4):
        o, g = ref[t], got[t]
        keep = (o["margin"] > 1e-7) | (o["status"] < 0)
        np.testing.assert_array_equal(g["status"][keep], o["status"][keep])
        np.testing.assert_array_equal(g["iters"][keep], o["iters"][keep])
        np.testing.assert_array_equal(np.isnan(g["torque"]), np.isnan(o["torque"]))
        fin = ~np.isnan(o["torque"]) & keep[:, None]
        d = np.abs(g["torque"] - o["torque"])[fin]
        assert (d <= (ATOL + RTOL * np.abs(o["torque"]))[fin]).all(), (preset, t, d.max())


@pytest.mark.parametrize("preset,config", [("walter_sr", "tumbling"), ("unitree_go2", "go2_standing")])
def test_dual_infeasibility_certificate_matches_oracle(oracle, host_core, preset, config):
    """OSQP's dual infeasibility certificate (util.c is_dual_infeasible).  The controller's own
    linear cost always lies in the range of H, so its QP is never unbounded; the harness adds
    an offset to f on a dof that H, M and J leave free (regularisation 0) to get one, and the
    oracle solves the same matrices through orc_setup / orc_solve."""
    import dataclasses
    import osc_b200 as ob
    from osc_b200 import capi
    spec = dataclasses.replace(ob.load_preset(preset), w_reg=0.0)
    N = 4
    k = 1  # an unactuated dof
    inp = {kk: v.copy() for kk, v in ob.synth.make_inputs(spec, N, config, step=0).items()}
    inp["M"][:, k, :] = 0
    inp["M"][:, :, k] = 0
    inp["J"][:, :, k] = 0
    inp["C"][:, k] = 0
    df = np.zeros(spec.nv)
    df[k] = -1.0
    ref = []
    for e in range(N):
        H, f, A, lo, hi = oracle.build_qp(spec, *[np.ascontiguousarray(inp[kk][e]) for kk in FIELDS])
        f = f.copy()
        f[:spec.nv] += df
        ref.append(oracle.solve_qp(H, f, A, lo, hi,
                                   oracle.default_settings(eps_abs=1e-7, eps_rel=1e-7)))
    assert all(r["status"] == -4 for r in ref), [r["status"] for r in ref]
    # (tolerances tight enough that the certificate fires before the growing iterates pass
    # the relative residual test)
    st = capi.CSettings(0.1, 1e-6, 1.6, 1e-7, 1e-7, 5.0, 10, 1, 0, 4000, 25, 1)
    host_core.osc_core_host_set_f_offset(_p(df), spec.nv)
    try:
        got = run_host_core(host_core, spec, st, [inp], N)[0]
    finally:
        host_core.osc_core_host_set_f_offset(None, 0)
    np.testing.assert_array_equal(got["status"], [r["status"] for r in ref])
    np.testing.assert_array_equal(got["iters"], [r["iter"] for r in ref])
    assert np.isnan(got["torque"]).all() and np.isnan(got["x"]).all()


def run_host_condensed(L, spec, settings, steps, n_envs):
    from osc_b200 import capi
    cs = capi.c_spec(spec)
    SS = L.osc_condensed_host_state_size(C.byref(cs))
    assert SS == 2 * (spec.nu + 3 * spec.nc) + 4 * spec.nc + 2
    outs = [dict(torque=np.zeros((n_envs, spec.nu)), iters=np.zeros(n_envs, np.int32),
                 status=np.zeros(n_envs, np.int32), rho=np.zeros(n_envs),
                 x=np.zeros((n_envs, spec.n)), y=np.zeros((n_envs, spec.m))) for _ in steps]
    ii = np.zeros(4, np.int32)
    dd = np.zeros(3)
    for e in range(n_envs):
        state = np.zeros(SS)
        x = np.zeros(spec.n); y = np.zeros(spec.m); tq = np.zeros(spec.nu)
        for t, data in enumerate(steps):
            a = [np.ascontiguousarray(data[k][e]) for k in FIELDS]
            L.osc_condensed_host_step(C.byref(cs), C.byref(settings), *[_p(v) for v in a],
                                      _p(state), _p(x), _p(y), _p(tq),
                                      ii.ctypes.data_as(C.POINTER(C.c_int)), _p(dd))
            o = outs[t]
            o["torque"][e], o["x"][e], o["y"][e] = tq, x, y
            o["iters"][e], o["status"][e], o["rho"][e] = ii[0], ii[1], dd[2]
    return outs


@pytest.mark.parametrize("preset,config", [("walter_sr_true_tumbling_mjjoint", "tumbling"),
                                           ("unitree_go2", "go2_standing"),
                                           ("walter_sr_wheels", "stairs")])
@pytest.mark.parametrize("kw", [dict(), dict(adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5),
                                dict(scaling=0, max_iter=150)],
                         ids=["defaults", "interval25-eps1e-5", "noscaling"])
def test_condensed_core_matches_condensed_oracle_on_host(oracle, host_core, preset, config, kw):
    """The CONDENSED fast mode's device source (csrc/osc_condensed.cuh: Cholesky of M,
    G = M^-1 [B Jc], P' = G'HdG + R, scaling, K^-1, ADMM) on the emulated warp vs its own
    oracle (oracle/osc_condensed.py: numpy condensation + the OSQP restatement's generic QP
    entry): iteration counts and status bit-exact, torques within 1e-5 + 1e-4 |tau|, cold step
    and two warm steps, both robot shapes."""
    import osc_b200 as ob
    import osc_condensed as oc
    from osc_b200 import capi
    spec = ob.load_preset(preset)
    N = 24
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(3)]
    ref = oc.CondensedOracle(spec, N, oracle.default_settings(**kw))
    got = run_host_condensed(host_core, spec, capi.default_settings(**kw) if False else
                             _csettings(kw), steps, N)
    for t, inp in enumerate(steps):
        o, g = ref.step(inp), got[t]
        np.testing.assert_array_equal(g["iters"], o["iters"])
        np.testing.assert_array_equal(g["status"], o["status"])
        d = np.abs(g["torque"] - o["torque"])
        tol = ATOL + RTOL * np.abs(o["torque"])
        assert (d <= tol).all(), (preset, t, (d / tol).max())
        np.testing.assert_allclose(g["rho"], o["rho"], rtol=1e-4)
        # solution in the reference's shape: [dv; u; z] with dv = G w + d0
        np.testing.assert_allclose(g["x"], o["x"], rtol=1e-3, atol=1e-3 * (1 + np.abs(o["x"]).max()))
        nv, nu, nc = spec.nv, spec.nu, spec.nc
        assert np.array_equal(g["torque"], g["x"][:, nv:nv + nu])
        # the eliminated dynamics hold exactly for the recovered dv
        Jc = inp["J"][:, 3 * spec.ns - 3 * nc:3 * spec.ns, :]
        dyn = (np.einsum("bij,bj->bi", inp["M"], g["x"][:, :nv]) + inp["C"]
               - np.concatenate([np.zeros((N, nv - nu)), g["x"][:, nv:nv + nu]], 1)
               - np.einsum("bki,bk->bi", Jc, g["x"][:, nv + nu:]))
        assert np.abs(dyn).max() < 1e-8 * (1 + np.abs(inp["C"]).max())


def _csettings(kw):
    """osc_settings without touching the CUDA library (the CPU suite has no GPU)."""
    from osc_b200 import capi
    s = capi.CSettings(0.1, 1e-6, 1.6, 1e-3, 1e-3, 5.0, 10, 1, 0, 4000, 25, 1)
    s.eps_prim_inf = s.eps_dual_inf = 1e-4
    for k, v in kw.items():
        setattr(s, k, v)
    return s


@pytest.mark.parametrize("preset,config", [("walter_sr_wheels", "stairs"), ("unitree_go2", "go2_standing")])
def test_condensed_core_has_no_cross_lane_hazard(host_core, host_core_reversed, preset, config):
    """The race check of test_no_cross_lane_hazard_inside_a_body for csrc/osc_condensed.cuh
    (in-place Cholesky, per-lane triangular solves, double-buffered exchanges, the two-pivot
    sweep on whole rows): lanes of every body in the opposite order, bit-identical results."""
    import osc_b200 as ob
    spec = ob.load_preset(preset)
    N = 8
    steps = [ob.synth.make_inputs(spec, N, config, step=t) for t in range(2)]
    st = _csettings(dict(adaptive_rho_interval=25))
    a = run_host_condensed(host_core, spec, st, steps, N)
    b = run_host_condensed(host_core_reversed, spec, st, steps, N)
    for t in range(2):
        for k in ("torque", "x", "y", "iters", "status", "rho"):
            assert np.array_equal(a[t][k], b[t][k]), (preset, t, k)
    assert (a[0]["rho"] != 0.1).any()
