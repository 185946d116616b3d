"""CPU: pins of oracle/osc_kinematics.py (the restatement of what the reference reads from
MuJoCo before its hot path, walter_sr/operational_space_controller.h:394-513) that do not
depend on its own recursions: finite differences of forward kinematics along trajectories
integrated with constant qvel (MuJoCo's conventions: world-frame base linear velocity,
body-frame base angular velocity)."""
import numpy as np
import pytest

import osc_kinematics as ok

H = 1e-6


def _vee(A):
    return np.array([A[2, 1] - A[1, 2], A[0, 2] - A[2, 0], A[1, 0] - A[0, 1]]) * 0.5


@pytest.mark.parametrize("make", [ok.walter_like_tree, ok.go2_like_tree])
def test_jacobians_bias_mass_and_bias_forces_against_finite_differences(make):
    tree = make(seed=3)
    qpos, qvel = ok.random_state(tree, 3, seed=5)
    A = tree.affects()
    for e in range(3):
        q, v = qpos[e], qvel[e]
        M, C, J, bias = ok.osc_data(tree, q, v)
        ns, nv = tree.ns, tree.nv
        assert J.shape == (6 * ns, nv) and M.shape == (nv, nv)
        # --- columns of J: unit velocity along every dof
        for d in range(nv):
            ed = np.zeros(nv); ed[d] = 1.0
            fp, fm = ok.forward(tree, ok.integrate(tree, q, ed, H)), ok.forward(tree, ok.integrate(tree, q, ed, -H))
            vp = (fp["site"] - fm["site"]) / (2 * H)
            np.testing.assert_allclose(J[:3 * ns, d].reshape(ns, 3), vp, atol=2e-8)
            for s in range(ns):
                b = tree.site_body[s]
                w = _vee(fp["R"][b] @ fm["R"][b].T) / H * 0.5 * 2
                np.testing.assert_allclose(J[3 * ns + 3 * s:3 * ns + 3 * s + 3, d], w / 2, atol=2e-8)
                if not A[b, d]:
                    assert not J[[3 * s, 3 * s + 1, 3 * s + 2], d].any()
        # --- bias = d/dt (J qvel) at qacc = 0
        qp, qm = ok.integrate(tree, q, v, H), ok.integrate(tree, q, v, -H)
        Jp, Jm = ok.osc_data(tree, qp, v)[2], ok.osc_data(tree, qm, v)[2]
        np.testing.assert_allclose(bias, (Jp @ v - Jm @ v) / (2 * H), rtol=1e-6, atol=1e-6)
        # --- M: kinetic energy from body velocities (numeric COM velocities, analytic omega)
        fk = ok.forward(tree, q, v)
        fp, fm = ok.forward(tree, qp), ok.forward(tree, qm)
        vc = (fp["com"] - fm["com"]) / (2 * H)
        T = 0.5 * sum(tree.mass[b] * vc[b] @ vc[b] + fk["w"][b] @ fk["Iw"][b] @ fk["w"][b]
                      for b in range(tree.nb))
        assert abs(0.5 * v @ M @ v - T) < 1e-6 * (1 + T)
        assert np.allclose(M, M.T, atol=1e-12) and np.linalg.eigvalsh(M).min() > 0
        # --- C: Newton-Euler with numerically differentiated body accelerations
        fkp, fkm = ok.forward(tree, qp, v), ok.forward(tree, qm, v)
        Cn = np.zeros(nv)
        for b in range(tree.nb):
            vcp = fkp["v"][b] + np.cross(fkp["w"][b], fkp["com"][b] - fkp["p"][b])
            vcm = fkm["v"][b] + np.cross(fkm["w"][b], fkm["com"][b] - fkm["p"][b])
            ac = (vcp - vcm) / (2 * H)
            al = (fkp["w"][b] - fkm["w"][b]) / (2 * H)
            jc, jr = ok.point_jacobian(tree, fk, b, fk["com"][b])
            Cn += jc.T @ (tree.mass[b] * (ac - ok.GRAVITY))
            Cn += jr.T @ (fk["Iw"][b] @ al + np.cross(fk["w"][b], fk["Iw"][b] @ fk["w"][b]))
        np.testing.assert_allclose(C, Cn, rtol=1e-5, atol=1e-5)


def test_free_fall_conserves_energy_and_momentum():
    """M qacc + C = 0 without gravity: total energy and linear momentum stay constant along
    an integrated trajectory (ties M and C together)."""
    tree = ok.walter_like_tree(seed=1)
    q, v = ok.random_state(tree, 1, seed=2)
    q, v = q[0], 0.5 * v[0]
    g0 = ok.GRAVITY.copy()
    ok.GRAVITY[:] = 0.0
    try:
        def energy_momentum(q, v):
            M = ok.osc_data(tree, q, v)[0]
            fk = ok.forward(tree, q, v)
            P = sum(tree.mass[b] * (fk["v"][b] + np.cross(fk["w"][b], fk["com"][b] - fk["p"][b]))
                    for b in range(tree.nb))
            return 0.5 * v @ M @ v, P
        E0, P0 = energy_momentum(q, v)
        h = 2e-5
        for _ in range(400):   # midpoint rule
            M, C, _, _ = ok.osc_data(tree, q, v)
            a1 = -np.linalg.solve(M, C)
            qh, vh = ok.integrate(tree, q, v, h / 2), v + 0.5 * h * a1
            M, C, _, _ = ok.osc_data(tree, qh, vh)
            a2 = -np.linalg.solve(M, C)
            q, v = ok.integrate(tree, q, vh, h), v + h * a2
        E1, P1 = energy_momentum(q, v)
        assert abs(E1 - E0) < 1e-6 * E0
        np.testing.assert_allclose(P1, P0, atol=1e-6 * (1 + np.abs(P0).max()))
    finally:
        ok.GRAVITY[:] = g0
