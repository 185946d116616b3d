"""oracle/osc_kinematics.py -- TEST INFRASTRUCTURE (CPU, numpy): the quantities the reference
reads from MuJoCo before its hot path -- update_mj_data / update_osc_data,
walter_sr/operational_space_controller.h:394-513 -- restated for a floating-base tree of hinge
joints:
    M     = mj_fullM(qM)                      (:436-438)   composite: sum_b m Jc'Jc + Jr' I Jr
    C     = qfrc_bias                         (:441-442)   gravity + Coriolis / centrifugal
    J     = [mj_jac  jacp ; jacr] per site    (:459-487)   rows [Jp(site 0..); Jr(site 0..)]
    bias  = mj_jacDot * qvel                  (:491-492)   = site acceleration at qacc = 0
with MuJoCo's conventions: qpos = [base xyz, base quat (w x y z), hinge angles], qvel = [base
linear velocity in the WORLD frame, base angular velocity in the BODY frame, hinge rates];
hinge anchors at the child body's origin.

MuJoCo and the robots' MJCF files are external to the reference and absent here, so parity
with MuJoCo itself is UNPINNED; what pins this restatement is independent of its own
recursions: finite differences of forward kinematics along integrated trajectories
(tests/test_kinematics_oracle.py).  The model constants come from a host struct (`Tree`) so the
real numbers can be dropped in when a user supplies them; `walter_like_tree` / `go2_like_tree`
build synthetic trees with the reference's topology (torso + 4 x (thigh, shin); base + 4 x
(hip, thigh, calf)) and site lists of the right sizes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

GRAVITY = np.array([0.0, 0.0, -9.81])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2])


def axis_angle_mat(axis, th):
    c, s = np.cos(th), np.sin(th)
    x, y, z = axis
    K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + s * K + (1 - c) * (K @ K)


@dataclass
class Tree:
    """Body 0 is the floating base; body b >= 1 hangs off parent[b] < b through one hinge
    about `jaxis[b]` (unit vector in the body's frame) anchored at the body's origin."""
    parent: np.ndarray        # [nb] int, parent[0] = -1
    bpos: np.ndarray          # [nb, 3] body origin in the parent frame (zero angle)
    bquat: np.ndarray         # [nb, 4] body orientation in the parent frame (zero angle)
    jaxis: np.ndarray         # [nb, 3]
    mass: np.ndarray          # [nb]
    ipos: np.ndarray          # [nb, 3] centre of mass in the body frame
    inertia: np.ndarray       # [nb, 3, 3] about the centre of mass, body frame
    site_body: np.ndarray     # [ns] int
    site_pos: np.ndarray      # [ns, 3] in the body frame
    nb: int = field(init=False)
    nv: int = field(init=False)
    nq: int = field(init=False)
    ns: int = field(init=False)

    def __post_init__(self):
        self.nb = len(self.parent)
        self.nv = 6 + self.nb - 1
        self.nq = 7 + self.nb - 1
        self.ns = len(self.site_body)
        assert self.parent[0] == -1 and all(self.parent[b] < b for b in range(1, self.nb))

    def affects(self):
        """[nb, nv] bool: dof d moves body b (free-joint dofs move everything)."""
        A = np.zeros((self.nb, self.nv), bool)
        A[:, :6] = True
        for b in range(1, self.nb):
            A[b] = A[self.parent[b]]
            A[b, 5 + b] = True
        return A

    def depth(self):
        d = np.zeros(self.nb, int)
        for b in range(1, self.nb):
            d[b] = d[self.parent[b]] + 1
        return d


def _rand_tree(rng, parent, n_sites_per_body):
    nb = len(parent)
    bpos = rng.uniform(-0.3, 0.3, (nb, 3)); bpos[0] = 0
    bquat = rng.normal(size=(nb, 4)); bquat /= np.linalg.norm(bquat, axis=1, keepdims=True); bquat[0] = [1, 0, 0, 0]
    jaxis = rng.normal(size=(nb, 3)); jaxis /= np.linalg.norm(jaxis, axis=1, keepdims=True)
    mass = rng.uniform(0.2, 4.0, nb); mass[0] = rng.uniform(5.0, 12.0)
    ipos = rng.uniform(-0.05, 0.05, (nb, 3))
    inertia = np.zeros((nb, 3, 3))
    for b in range(nb):
        A = rng.normal(size=(3, 3))
        inertia[b] = mass[b] * (0.01 * np.eye(3) + 0.005 * A @ A.T)
    site_body, site_pos = [], []
    for b, k in enumerate(n_sites_per_body):
        for _ in range(k):
            site_body.append(b)
            site_pos.append(rng.uniform(-0.2, 0.2, 3))
    return Tree(np.array(parent), bpos, bquat, jaxis, mass, ipos, inertia, np.array(site_body),
                np.array(site_pos))


def walter_like_tree(seed=0):
    """Walter Sr's topology: torso + 4 x (thigh -> shin), nv = 14; 17 sites in the controller's
    order [torso, 4 shins, 4 thighs, 8 wheel contacts (two per shin)]
    (config/walter_sr/walter_sr_config.yaml:20-64)."""
    rng = np.random.default_rng([seed, 14])
    parent = [-1, 0, 1, 0, 3, 0, 5, 0, 7]  # thigh / shin pairs
    t = _rand_tree(rng, parent, [0] * 9)
    shins, thighs = [2, 4, 6, 8], [1, 3, 5, 7]
    body = [0] + shins + thighs + [s for s in shins for _ in range(2)]
    t.site_body = np.array(body)
    t.site_pos = rng.uniform(-0.2, 0.2, (17, 3))
    t.ns = 17
    return t


def go2_like_tree(seed=0):
    """Unitree Go2's topology: base + 4 x (hip -> thigh -> calf), nv = 18; sites [base, 4 feet]
    (config/unitree_go2/unitree_go2_config.yaml:8-15)."""
    rng = np.random.default_rng([seed, 18])
    parent = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11]
    t = _rand_tree(rng, parent, [0] * 13)
    t.site_body = np.array([0, 3, 6, 9, 12])
    t.site_pos = rng.uniform(-0.2, 0.2, (5, 3))
    t.ns = 5
    return t


def integrate(tree: Tree, qpos, qvel, h):
    """mj_integratePos with constant qvel for a time h."""
    q = np.array(qpos, float)
    q[0:3] += h * qvel[0:3]
    w = np.asarray(qvel[3:6], float)
    n = np.linalg.norm(w)
    dq = np.array([1.0, 0, 0, 0]) if n * abs(h) < 1e-300 else np.concatenate(
        [[np.cos(0.5 * n * h)], np.sin(0.5 * n * h) * w / n])
    q[3:7] = quat_mul(q[3:7], dq)          # angular velocity is expressed in the BODY frame
    q[3:7] /= np.linalg.norm(q[3:7])
    q[7:] += h * qvel[6:]
    return q


def forward(tree: Tree, qpos, qvel=None):
    """Frames, and with qvel the velocities and the accelerations at qacc = 0, per body."""
    nb = tree.nb
    R = np.zeros((nb, 3, 3)); p = np.zeros((nb, 3))
    R[0] = quat_to_mat(qpos[3:7] / np.linalg.norm(qpos[3:7])); p[0] = qpos[0:3]
    for b in range(1, nb):
        pa = tree.parent[b]
        R[b] = R[pa] @ quat_to_mat(tree.bquat[b]) @ axis_angle_mat(tree.jaxis[b], qpos[6 + b])
        p[b] = p[pa] + R[pa] @ tree.bpos[b]
    axis = np.zeros((tree.nv, 3)); anchor = np.zeros((tree.nv, 3))
    axis[0:3] = np.eye(3)
    axis[3:6] = R[0].T                     # rows = body axes in the world frame
    anchor[3:6] = p[0]
    for b in range(1, nb):
        axis[5 + b] = R[b] @ tree.jaxis[b]
        anchor[5 + b] = p[b]
    out = dict(R=R, p=p, axis=axis, anchor=anchor,
               com=p + np.einsum("bij,bj->bi", R, tree.ipos),
               Iw=np.einsum("bij,bjk,blk->bil", R, tree.inertia, R),
               site=p[tree.site_body] + np.einsum("sij,sj->si", R[tree.site_body], tree.site_pos))
    if qvel is None:
        return out
    w = np.zeros((nb, 3)); v = np.zeros((nb, 3)); al = np.zeros((nb, 3)); a = np.zeros((nb, 3))
    w[0] = R[0] @ qvel[3:6]; v[0] = qvel[0:3]
    for b in range(1, nb):
        pa = tree.parent[b]
        r = p[b] - p[pa]
        wj = axis[5 + b] * qvel[5 + b]
        w[b] = w[pa] + wj
        v[b] = v[pa] + np.cross(w[pa], r)
        al[b] = al[pa] + np.cross(w[pa], wj)
        a[b] = a[pa] + np.cross(al[pa], r) + np.cross(w[pa], np.cross(w[pa], r))
    out.update(w=w, v=v, al=al, a=a)
    return out


def point_jacobian(tree: Tree, fk, body, x):
    """mj_jac: (jacp, jacr), each [3, nv], of world point x attached to `body`."""
    A = tree.affects()[body]
    jp = np.zeros((3, tree.nv)); jr = np.zeros((3, tree.nv))
    for d in range(tree.nv):
        if not A[d]:
            continue
        if d < 3:
            jp[:, d] = fk["axis"][d]
        else:
            jp[:, d] = np.cross(fk["axis"][d], x - fk["anchor"][d])
            jr[:, d] = fk["axis"][d]
    return jp, jr


def osc_data(tree: Tree, qpos, qvel):
    """M [nv, nv], C [nv], J [6 ns, nv], bias [6 ns] of one environment."""
    fk = forward(tree, qpos, qvel)
    nv, ns = tree.nv, tree.ns
    M = np.zeros((nv, nv)); C = np.zeros(nv)
    for b in range(tree.nb):
        jc, jr = point_jacobian(tree, fk, b, fk["com"][b])
        Iw = fk["Iw"][b]
        M += tree.mass[b] * jc.T @ jc + jr.T @ Iw @ jr
        rho = fk["com"][b] - fk["p"][b]
        ac = fk["a"][b] + np.cross(fk["al"][b], rho) + np.cross(fk["w"][b], np.cross(fk["w"][b], rho))
        F = tree.mass[b] * (ac - GRAVITY)
        N = Iw @ fk["al"][b] + np.cross(fk["w"][b], Iw @ fk["w"][b])
        C += jc.T @ F + jr.T @ N
    J = np.zeros((6 * ns, nv)); bias = np.zeros(6 * ns)
    for s in range(ns):
        b = tree.site_body[s]
        jp, jr = point_jacobian(tree, fk, b, fk["site"][s])
        J[3 * s:3 * s + 3] = jp
        J[3 * ns + 3 * s:3 * ns + 3 * s + 3] = jr
        rho = fk["site"][s] - fk["p"][b]
        bias[3 * s:3 * s + 3] = (fk["a"][b] + np.cross(fk["al"][b], rho)
                                 + np.cross(fk["w"][b], np.cross(fk["w"][b], rho)))
        bias[3 * ns + 3 * s:3 * ns + 3 * s + 3] = fk["al"][b]
    return M, C, J, bias


def osc_data_batch(tree: Tree, qpos, qvel):
    out = [osc_data(tree, qpos[e], qvel[e]) for e in range(len(qpos))]
    return tuple(np.stack([o[k] for o in out]) for k in range(4))


def random_state(tree: Tree, n, seed=0, first_env=0):
    """qpos [n, nq] (base at the origin like update_mj_data :402-403, unit quaternion), qvel [n, nv]."""
    rng = np.random.default_rng([seed, first_env, 77])
    qpos = np.zeros((n, tree.nq)); qvel = rng.normal(0.0, 1.0, (n, tree.nv))
    q = rng.normal(size=(n, 4))
    qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qpos[:, 7:] = rng.uniform(-1.2, 1.2, (n, tree.nq - 7))
    return qpos, qvel
