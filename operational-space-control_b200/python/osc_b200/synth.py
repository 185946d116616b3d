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
                first_env: int = 0):
    """Returns dict of float64 arrays M[N,nv,nv], C[N,nv], J[N,s,nv], bias[N,s],
    targets[N,ns,6], mask[N,nc].  Environment e's data depends only on
    (seed, first_env + e), so shards of one job agree with the unsharded job."""
    nv, ns, nc, s = spec.nv, spec.ns, spec.nc, spec.s
    N = n_envs
    out = {k: None for k in ("M", "C", "J", "bias", "targets", "mask")}
    M = np.empty((N, nv, nv)); C = np.empty((N, nv)); J = np.empty((N, s, nv))
    bias = np.empty((N, s)); targets = np.zeros((N, ns, 6)); mask = np.ones((N, nc))
    # per-environment streams in blocks (keeps generation vectorised and shard-invariant)
    BLK = 256
    for b0 in range(0, N, BLK):
        b1 = min(N, b0 + BLK)
        g0 = first_env + b0
        assert g0 % BLK == 0 or N <= BLK or first_env % BLK == 0, "shards must start on 256-env blocks"
        rng = np.random.default_rng([seed, g0 // BLK])
        nb_full = BLK
        Jb = rng.normal(0.0, 0.3, size=(nb_full, s, nv))
        for i in range(ns):
            Jb[:, 3 * i:3 * i + 3, 0:3] = np.eye(3)
            Jb[:, 3 * ns + 3 * i:3 * ns + 3 * i + 3, 0:3] = 0.0
            Jb[:, 3 * ns + 3 * i:3 * ns + 3 * i + 3, 3:6] = np.eye(3)
        mass = rng.uniform(0.05, 1.5, size=(nb_full, ns))
        inert = rng.uniform(1e-3, 2e-2, size=(nb_full, ns))
        Jp = Jb[:, :3 * ns].reshape(nb_full, ns, 3, nv)
        Jr = Jb[:, 3 * ns:].reshape(nb_full, ns, 3, nv)
        Mb = (np.einsum("bi,bika,bikc->bac", mass, Jp, Jp)
              + np.einsum("bi,bika,bikc->bac", inert, Jr, Jr) + 1e-3 * np.eye(nv))
        Mb = 0.5 * (Mb + Mb.transpose(0, 2, 1))
        grav = np.array([0.0, 0.0, -9.81])
        Cb = -np.einsum("bi,bika,k->ba", mass, Jp, grav) + rng.normal(0.0, 0.5, size=(nb_full, nv))
        biasb = rng.normal(0.0, 1.0, size=(nb_full, s))
        tb = np.zeros((nb_full, ns, 6)); mb = np.ones((nb_full, nc))
        if config == "standing":        # walter_sr_standing.cc:141-146: zero targets, all contacts
            pass
        elif config == "go2_standing":  # standing.cc:146-155: base PD row only
            tb[:, 0, :] = rng.uniform(-5.0, 5.0, size=(nb_full, 6))
        elif config in ("tumbling", "stairs"):
            # walter_sr_true_tumbling_mjjoint.cc:756-781 (shin rows 1-4, alpha_y),
            # :873-946 (thigh rows 5-8, z) ; row 0 zero (:1001-1019)
            if ns >= 9:
                tb[:, 1:5, 4] = rng.uniform(-2400.0, 2400.0, size=(nb_full, 4))
                tb[:, 5:9, 2] = rng.uniform(-200.0, 200.0, size=(nb_full, 4))
            if config == "stairs":
                # walter_sr_true_stairclimbing_mjjoint.cc:997-1012 torso row
                tb[:, 0, 0] = rng.uniform(-20.0, 20.0, size=nb_full)
                tb[:, 0, 3:6] = rng.uniform(-5.0, 5.0, size=(nb_full, 3))
                mb = (rng.uniform(size=(nb_full, nc)) < 0.75).astype(np.float64)
            else:
                mb = (rng.uniform(size=(nb_full, nc)) < 0.6).astype(np.float64)
        else:
            raise ValueError(config)
        k = b1 - b0
        M[b0:b1] = Mb[:k]; C[b0:b1] = Cb[:k]; J[b0:b1] = Jb[:k]; bias[b0:b1] = biasb[:k]
        targets[b0:b1] = tb[:k]; mask[b0:b1] = mb[:k]
    out.update(M=M, C=C, J=J, bias=bias, targets=targets, mask=mask)
    return out


def perturb(inputs: dict, rel: float = 0.01, seed: int = 1):
    """Next control step's inputs: the same environments 'a little later'
    (warm = previous solution of a perturbed problem)."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, v in inputs.items():
        if k == "mask":
            out[k] = v.copy()
        elif k == "M":
            # keep M symmetric positive definite: perturb congruently
            N, nv, _ = v.shape
            S = np.eye(nv) + rel * rng.normal(size=(N, nv, nv))
            out[k] = np.einsum("bij,bjk,blk->bil", S, v, S)
            out[k] = 0.5 * (out[k] + out[k].transpose(0, 2, 1))
        elif k == "J":
            Jn = v * (1.0 + rel * rng.normal(size=v.shape))
            out[k] = np.where((v == 0.0) | (v == 1.0), v, Jn)  # keep floating-base structure
        else:
            out[k] = v * (1.0 + rel * rng.normal(size=v.shape))
    return out
