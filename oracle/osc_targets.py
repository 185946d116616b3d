"""CPU restatement of the step BEFORE the hot path: task-space PD
targets and the contact mask, as the reference's example drivers compute them.

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/ may import this module.
PARITY UNPINNED by the reference (it ships no tests); pinned here by hand-worked values in
tests/test_targets_oracle.py.

Follows (paths relative to /root/reference):
  examples/standing.cc:146-155                       the PD law on one site
  examples/walter_sr_true_tumbling_mjjoint.cc:42     contains()
  examples/walter_sr_true_tumbling_mjjoint.cc:106    getSiteIdsOnSameBodyAsGeom()
  examples/walter_sr_true_tumbling_mjjoint.cc:152    getBinaryRepresentation_std_find()
  examples/walter_sr_true_tumbling_mjjoint.cc:523-558  contacts -> contact_mask
  examples/walter_sr_true_tumbling_mjjoint.cc:695-802  shin rows 1-4 (alpha_y from the joint angle)
  examples/walter_sr_true_tumbling_mjjoint.cc:873-973  thigh rows 5-8 (height law), row 0 (:1001-1019)
"""
from __future__ import annotations

import numpy as np


def quat_mul(a, b):
    """Eigen::Quaternion product, (w, x, y, z) order, broadcasting over leading axes."""
    aw, ax, ay, az = (a[..., i] for i in range(4))
    bw, bx, by, bz = (b[..., i] for i in range(4))
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ], axis=-1)


def quat_conj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def targets_pd(pos, quat, vel, angvel, pos_des, quat_des, kp_lin, kd_lin, kp_ang, kd_ang,
               vel_des=None, angvel_des=None):
    """standing.cc:146-155 for every (environment, site):
         linear  = kp_lin (p_des - p) + kd_lin (v_des - v)
         angular = kp_ang vec(q_des * conj(q)) + kd_ang (w_des - w)
       Shapes [n_envs, ns, 3|4]; gains [ns].  Returns TaskspaceTargets [n_envs, ns, 6]."""
    pos, quat, vel, angvel = (np.asarray(a, dtype=np.float64) for a in (pos, quat, vel, angvel))
    vel_des = np.zeros_like(vel) if vel_des is None else np.asarray(vel_des, dtype=np.float64)
    angvel_des = (np.zeros_like(angvel) if angvel_des is None
                  else np.asarray(angvel_des, dtype=np.float64))
    kp_lin, kd_lin, kp_ang, kd_ang = (np.asarray(g, dtype=np.float64)[None, :, None]
                                      for g in (kp_lin, kd_lin, kp_ang, kd_ang))
    rot_err = quat_mul(np.asarray(quat_des, dtype=np.float64), quat_conj(quat))[..., 1:]
    lin = kp_lin * (np.asarray(pos_des, dtype=np.float64) - pos) + kd_lin * (vel_des - vel)
    ang = kp_ang * rot_err + kd_ang * (angvel_des - angvel)
    return np.concatenate([lin, ang], axis=-1)


def contact_mask_from_contacts(geom_pairs, ncon, listed, site_of_geom=None):
    """walter_sr_true_tumbling_mjjoint.cc:523-558, literal loops (small cases only).
    geom_pairs [n_envs, max_con, 2] int, ncon [n_envs] int, listed = wheel_sites_mujoco,
    site_of_geom[j] = site the j-th listed geom maps to (None: ids coincide)."""
    listed = [int(v) for v in listed]
    site_of = dict(zip(listed, listed if site_of_geom is None else [int(v) for v in site_of_geom]))
    n_envs = len(ncon)
    mask = np.zeros((n_envs, len(listed)))
    for e in range(n_envs):
        sites = []
        for slot in (1, 0):  # the reference scans geom[1] first, then geom[0]
            for k in range(int(ncon[e])):
                g = int(geom_pairs[e, k, slot])
                if g in site_of:
                    sites.append(site_of[g])
        for c, s in enumerate(listed):
            mask[e, c] = 1.0 if s in sites else 0.0
    return mask


# the numbers of examples/walter_sr_true_tumbling_mjjoint.cc (BASELINE.json configs[2])
WALTER_TUMBLING = dict(shin_kp=800.0 * 3.0, shin_kv=800.0 * 3.0, shin_rate=0.1 * 8.0 * 5.0,
                       thigh_kp=4000.0 * 0.5, thigh_kv=600.0 * 0.5, thigh_rate=0.0,
                       thigh_height_offset=-0.025)


def targets_walter_tumbling(shin_angle, shin_angle_prev, shin_angle0, thigh_z, thigh_z_prev,
                            thigh_z0, time, dt, ns=17, **kw):
    """walter_sr_true_tumbling_mjjoint.cc per environment.  Arrays [N, 4] in the driver's leg
    order (tl, tr, hl, hr); returns targets [N, ns, 6]:
      rows 1-4 (shins, :695-802):  (0,0,0, 0, kp (th0 + rate t - th) + kv (rate - (th - th_prev)/dt), 0)
      rows 5-8 (thighs, :873-973): (0,0, kp ((z0 + offset) - z) + kv (rate_z - (z - z_prev)/dt), 0,0,0)
      row 0 (torso, :1001-1019): all gains zero -> zeros; contact rows 9-16: never written."""
    g = dict(WALTER_TUMBLING, **kw)
    sa, sp, s0 = (np.asarray(a, float) for a in (shin_angle, shin_angle_prev, shin_angle0))
    tz, tp, t0 = (np.asarray(a, float) for a in (thigh_z, thigh_z_prev, thigh_z0))
    N = sa.shape[0]
    out = np.zeros((N, ns, 6))
    shin_vel = (sa - sp) / dt
    out[:, 1:5, 4] = g["shin_kp"] * ((s0 + g["shin_rate"] * time) - sa) + g["shin_kv"] * (g["shin_rate"] - shin_vel)
    thigh_vel = (tz - tp) / dt
    out[:, 5:9, 2] = (g["thigh_kp"] * ((t0 - 0.0 + g["thigh_height_offset"]) - tz)
                      + g["thigh_kv"] * (g["thigh_rate"] - thigh_vel))
    return out
