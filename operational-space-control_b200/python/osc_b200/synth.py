"""Synthetic batched controller inputs (SURVEY.md 8(d) "Value distributions").

MuJoCo and the robot MJCF models are external to the reference and absent
here, so the quantities the reference reads from MuJoCo (`update_osc_data`,
`walter_sr/operational_space_controller.h:434-513`) are drawn with the
structure a floating-base robot gives them:
  J    : translational rows cols 0-2 = I3, rotational rows cols 0-2 = 0 and
         cols 3-5 = I3, everything else N(0, 0.3^2);
  M    : eps I + sum_i m_i Jp_i' Jp_i + Jr_i' I_i Jr_i  (SPD, consistent with J);
  C    : -sum_i m_i Jp_i' g + N(0, 0.5^2);   bias : N(0, 1).
Targets and contact masks follow the example drivers named by each
BASELINE.json config.
"""
from __future__ import annotations

import numpy as np

from .specs import RobotSpec

CONFIGS = ("standing", "go2_standing", "tumbling", "stairs")


def make_inputs(spec: RobotSpec, n_envs: int, config: str = "tumbling", seed: int = 0xB200,
                first_env: int = 0, step: int = 0, rel: float = 0.01):
    """Returns dict of float64 arrays M[N,nv,nv], C[N,nv], J[N,s,nv], bias[N,s],
    targets[N,ns,6], mask[N,nc].

    Environment e's data depends only on (seed, first_env + e, step), so shards of one
    job agree with the unsharded job (shards must start on multiples of 256 envs).
    step > 0 gives "the same environments a little later": every non-structural
    quantity of step 0 moved by a relative `rel` (1 %), M rebuilt from the moved J so
    that it keeps the floating-base structure (exact zeros stay exact zeros, which is
    what keeps OSQP on its same-sparsity update path, reference :565-570); contact
    masks are kept."""
    nv, ns, nc, s = spec.nv, spec.ns, spec.nc, spec.s
    N = n_envs
    BLK = 256
    assert first_env % BLK == 0, "shards must start on 256-env blocks"
    M = np.empty((N, nv, nv)); C = np.empty((N, nv)); J = np.empty((N, s, nv))
    bias = np.empty((N, s)); targets = np.zeros((N, ns, 6)); mask = np.ones((N, nc))
    grav = np.array([0.0, 0.0, -9.81])
    for b0 in range(0, N, BLK):
        b1 = min(N, b0 + BLK)
        blk = (first_env + b0) // BLK
        rng = np.random.default_rng([seed, blk])
        Jb = rng.normal(0.0, 0.3, size=(BLK, s, nv))
        mass = rng.uniform(0.05, 1.5, size=(BLK, ns))
        inert = rng.uniform(1e-3, 2e-2, size=(BLK, ns))
        Cn = rng.normal(0.0, 0.5, size=(BLK, nv))
        biasb = rng.normal(0.0, 1.0, size=(BLK, s))
        tb = np.zeros((BLK, ns, 6)); mb = np.ones((BLK, nc))
        if config == "standing":        # walter_sr_standing.cc:141-146: zero targets, all contacts
            pass
        elif config == "go2_standing":  # standing.cc:146-155: base PD row only
            tb[:, 0, :] = rng.uniform(-5.0, 5.0, size=(BLK, 6))
        elif config in ("tumbling", "stairs"):
            # walter_sr_true_tumbling_mjjoint.cc:756-781 (shin rows 1-4, alpha_y),
            # :873-946 (thigh rows 5-8, z) ; row 0 zero (:1001-1019)
            if ns >= 9:
                tb[:, 1:5, 4] = rng.uniform(-2400.0, 2400.0, size=(BLK, 4))
                tb[:, 5:9, 2] = rng.uniform(-200.0, 200.0, size=(BLK, 4))
            if config == "stairs":
                # walter_sr_true_stairclimbing_mjjoint.cc:997-1012 torso row
                tb[:, 0, 0] = rng.uniform(-20.0, 20.0, size=BLK)
                tb[:, 0, 3:6] = rng.uniform(-5.0, 5.0, size=(BLK, 3))
                mb = (rng.uniform(size=(BLK, nc)) < 0.75).astype(np.float64)
            else:
                mb = (rng.uniform(size=(BLK, nc)) < 0.6).astype(np.float64)
        else:
            raise ValueError(config)
        if step > 0:
            rs = np.random.default_rng([seed, blk, step])
            Jb = Jb * (1.0 + rel * rs.normal(size=Jb.shape))
            Cn = Cn * (1.0 + rel * rs.normal(size=Cn.shape))
            biasb = biasb * (1.0 + rel * rs.normal(size=biasb.shape))
            tb = tb * (1.0 + rel * rs.normal(size=tb.shape))
            mass = mass * (1.0 + rel * rs.normal(size=mass.shape))
        for i in range(ns):
            Jb[:, 3 * i:3 * i + 3, 0:3] = np.eye(3)
            Jb[:, 3 * ns + 3 * i:3 * ns + 3 * i + 3, 0:3] = 0.0
            Jb[:, 3 * ns + 3 * i:3 * ns + 3 * i + 3, 3:6] = np.eye(3)
        Jp = Jb[:, :3 * ns].reshape(BLK, ns, 3, nv)
        Jr = Jb[:, 3 * ns:].reshape(BLK, ns, 3, nv)
        Mb = (np.einsum("bi,bika,bikc->bac", mass, Jp, Jp)
              + np.einsum("bi,bika,bikc->bac", inert, Jr, Jr) + 1e-3 * np.eye(nv))
        Mb = 0.5 * (Mb + Mb.transpose(0, 2, 1))
        Cb = -np.einsum("bi,bika,k->ba", mass, Jp, grav) + Cn
        k = b1 - b0
        M[b0:b1] = Mb[:k]; C[b0:b1] = Cb[:k]; J[b0:b1] = Jb[:k]; bias[b0:b1] = biasb[:k]
        targets[b0:b1] = tb[:k]; mask[b0:b1] = mb[:k]
    return dict(M=M, C=C, J=J, bias=bias, targets=targets, mask=mask)
