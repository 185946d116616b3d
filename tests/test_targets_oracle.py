"""CPU: the restatement of the step before the hot path (oracle/osc_targets.py) against
hand-worked values of the reference's formulas (examples/standing.cc:146-155,
examples/walter_sr_true_tumbling_mjjoint.cc:523-558).  The reference ships no fixtures for
these; the values below were worked out by hand from the cited lines."""
import numpy as np

import osc_targets as ot


def test_quaternion_product_is_eigens():
    # i * j = k, j * i = -k (Hamilton convention, what Eigen implements)
    i, j = np.array([0.0, 1, 0, 0]), np.array([0.0, 0, 1, 0])
    assert np.array_equal(ot.quat_mul(i, j), [0, 0, 0, 1])
    assert np.array_equal(ot.quat_mul(j, i), [0, 0, 0, -1])
    q = np.array([0.5, -0.5, 0.5, 0.5])
    assert np.allclose(ot.quat_mul(q, ot.quat_conj(q)), [1, 0, 0, 0])


def test_standing_pd_law_by_hand():
    # standing.cc:146-155 with gains 150/25/50/10: body 0.1 m below its initial height,
    # sinking at 0.2 m/s, rolled by 0.2 rad about x, rolling at 0.3 rad/s
    th = 0.2
    q = np.array([np.cos(th / 2), np.sin(th / 2), 0, 0])
    t = ot.targets_pd(pos=[[[0, 0, 0.2]]], quat=[[q]], vel=[[[0, 0, -0.2]]],
                      angvel=[[[0.3, 0, 0]]], pos_des=[[[0, 0, 0.3]]],
                      quat_des=[[[1.0, 0, 0, 0]]], kp_lin=[150.0], kd_lin=[25.0],
                      kp_ang=[50.0], kd_ang=[10.0])
    # linear: 150*0.1 + 25*0.2 = 20 on z; angular: 50*(-sin(0.1)) + 10*(-0.3) on x
    expect = [0, 0, 20.0, 50.0 * -np.sin(0.1) - 3.0, 0, 0]
    np.testing.assert_allclose(t[0, 0], expect, rtol=0, atol=1e-14)


def test_pd_broadcasts_over_envs_and_sites():
    rng = np.random.default_rng(0)
    n, ns = 5, 3
    a = {k: rng.standard_normal((n, ns, 3)) for k in ("pos", "vel", "angvel", "pos_des")}
    quat = rng.standard_normal((n, ns, 4))
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    qd = rng.standard_normal((n, ns, 4))
    g = rng.uniform(1, 100, (4, ns))
    t = ot.targets_pd(quat=quat, quat_des=qd, kp_lin=g[0], kd_lin=g[1], kp_ang=g[2],
                      kd_ang=g[3], **a)
    for e in range(n):
        for s in range(ns):
            one = ot.targets_pd(pos=a["pos"][e:e + 1, s:s + 1], quat=quat[e:e + 1, s:s + 1],
                                vel=a["vel"][e:e + 1, s:s + 1],
                                angvel=a["angvel"][e:e + 1, s:s + 1],
                                pos_des=a["pos_des"][e:e + 1, s:s + 1],
                                quat_des=qd[e:e + 1, s:s + 1], kp_lin=g[0, s:s + 1],
                                kd_lin=g[1, s:s + 1], kp_ang=g[2, s:s + 1], kd_ang=g[3, s:s + 1])
            assert np.array_equal(one[0, 0], t[e, s])


def test_contact_mask_by_hand():
    listed = [3, 4, 7, 8, 11, 12, 15, 16]  # wheel_sites_mujoco (:436)
    pairs = np.zeros((3, 4, 2), np.int32)
    ncon = np.array([3, 0, 4], np.int32)
    # env 0: floor(0)-geom 4, floor-geom 11, geom 16 in slot 0; a 4th (stale) entry is ignored
    pairs[0] = [[0, 4], [0, 11], [16, 0], [0, 3]]
    # env 2: non-listed geoms and a duplicate
    pairs[2] = [[0, 5], [7, 7], [0, 7], [2, 9]]
    m = ot.contact_mask_from_contacts(pairs, ncon, listed)
    assert np.array_equal(m[0], [0, 1, 0, 0, 1, 0, 0, 1])
    assert np.array_equal(m[1], np.zeros(8))
    assert np.array_equal(m[2], [0, 0, 1, 0, 0, 0, 0, 0])
    # a geom -> site table that is not the identity: geom 4's body carries site 3
    m2 = ot.contact_mask_from_contacts(pairs, ncon, listed, [3, 3, 7, 8, 11, 12, 15, 16])
    assert np.array_equal(m2[0], [1, 0, 0, 0, 1, 0, 0, 1])


def test_walter_tumbling_laws_hand_worked():
    """walter_sr_true_tumbling_mjjoint.cc:695-802 (shins), :873-973 (thighs): one environment by
    hand with the driver's gains (2400 / 2400 / 4 rad/s; 2000 / 300; -0.025 m)."""
    import osc_targets as ot
    t = ot.targets_walter_tumbling(shin_angle=[[0.5, 0, 0, 0]], shin_angle_prev=[[0.498, 0, 0, 0]],
                                   shin_angle0=[[0.1, 0, 0, 0]], thigh_z=[[0, 0.30, 0, 0]],
                                   thigh_z_prev=[[0, 0.301, 0, 0]], thigh_z0=[[0, 0.33, 0, 0]],
                                   time=0.1, dt=0.002)
    assert t.shape == (1, 17, 6)
    # shin tl: 2400 ((0.1 + 4*0.1) - 0.5) + 2400 (4 - 0.002/0.002) = 0 + 2400*3
    assert abs(t[0, 1, 4] - 7200.0) < 1e-9
    # thigh tr: 2000 ((0.33 - 0.025) - 0.30) + 300 (0 - (-0.001/0.002)) = 10 + 150
    assert abs(t[0, 6, 2] - 160.0) < 1e-9
    nz = np.zeros((17, 6), bool)
    nz[1:5, 4] = True
    nz[5:9, 2] = True
    assert not t[0][~nz].any()   # torso row and contact rows stay zero
