"""Robot descriptors for the batched operational-space controller.

A `RobotSpec` carries exactly what the reference bakes into its build for one
robot: model sizes (`autogen_defines.h`, reference `<robot>/autogen/autogen.py:469-513`),
the objective weights and friction coefficient of the config YAML
(`config/<robot>/*.yaml`, consumed unchanged), and the hard-coded torque /
contact-force bounds (`walter_sr/operational_space_controller.h:284-353`,
`unitree_go2/operational_space_controller.h:285-308`).
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import List

_PRESET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "presets")


@dataclasses.dataclass(frozen=True)
class RobotSpec:
    name: str
    robot: str
    nq: int
    nv: int
    nu: int
    ns: int  # task sites (non-contact first, then contact)
    nc: int  # contact sites
    mu: float
    w_trans: tuple
    w_rot: tuple
    w_torque: float
    w_reg: float
    u_lb: tuple
    u_ub: tuple
    fz_max: float

    # sizes, named like the reference's constants.h / autogen_defines.h
    @property
    def nz(self) -> int:  # z_size
        return 3 * self.nc

    @property
    def n(self) -> int:  # design_vector_size
        return self.nv + self.nu + self.nz

    @property
    def m(self) -> int:  # constraint_matrix_rows
        return self.nv + 4 * self.nc + self.n

    @property
    def s(self) -> int:  # s_size
        return 6 * self.ns

    @property
    def in_doubles(self) -> int:
        """FP64 scalars crossing the boundary per solve (M, C, J, bias, targets, mask)."""
        return self.nv * self.nv + self.nv + self.s * self.nv + self.s + self.s + self.nc

    @property
    def algorithmic_bytes(self) -> int:
        """compulsory bytes per solve, inputs + torque, no warm start."""
        return 8 * (self.in_doubles + self.nu)


def preset_names() -> List[str]:
    return sorted(f[:-5] for f in os.listdir(_PRESET_DIR) if f.endswith(".json"))


def _from_dict(name: str, d: dict) -> RobotSpec:
    return RobotSpec(
        name=name, robot=d["robot"], nq=d["nq"], nv=d["nv"], nu=d["nu"], ns=d["ns"], nc=d["nc"],
        mu=float(d["mu"]), w_trans=tuple(d["w_trans"]), w_rot=tuple(d["w_rot"]),
        w_torque=float(d["w_torque"]), w_reg=float(d["w_reg"]),
        u_lb=tuple(d["u_lb"]), u_ub=tuple(d["u_ub"]), fz_max=float(d["fz_max"]))


def load_preset(name: str) -> RobotSpec:
    with open(os.path.join(_PRESET_DIR, f"{name}.json")) as fh:
        return _from_dict(name, json.load(fh))


def from_yaml(path: str, robot: str) -> RobotSpec:
    """Build a spec straight from a reference-format config YAML."""
    import importlib.util
    tool = os.path.join(os.path.dirname(_PRESET_DIR), "..", "tools", "gen_presets.py")
    spec = importlib.util.spec_from_file_location("gen_presets", tool)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return _from_dict(os.path.basename(path), mod.preset_from_yaml(robot, path))
