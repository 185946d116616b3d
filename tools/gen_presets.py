#!/usr/bin/env python3
"""Generate robot presets (JSON) and the reference-named `autogen_defines.h`
headers from the reference's *unchanged* config YAML files.

The reference turns its YAML into compile-time constants with a CasADi/MuJoCo
build step (`<robot>/autogen/autogen.py:469-518`, Bazel genrule
`<robot>/autogen/BUILD.bazel:20-29`).  This tool is the replacement for that
step: same YAML in, same constant names out, no CasADi (the QP matrices are
closed forms) and no MuJoCo (nq/nv/nu come from the table
below because the MJCF models are an external, un-vendored dependency --
SURVEY.md section 8 "sizes").

It also emits `autogen_functions.{cc,h}` with the C interface of the CasADi-generated
file the reference compiles (functions `beq Aeq bineq Aineq H f`, each with
`_incref _decref _checkout _release`, the `*_SZ_ARG/RES/IW/W` macros and the
`casadi_int` / `casadi_real` typedefs -- what walter_sr/utilities.h:10-77 and
operational_space_controller.h:43-106 bind), implemented from the closed forms of
autogen.py:54-345 for ANY design vector q (the reference always passes q = 0).  With
those three files the Bazel genrule `<robot>/autogen/BUILD.bazel:20-29`
(`cmd = "$(location :autogen) --filepath=$(RULEDIR)"`) can call this script instead of
the CasADi/MuJoCo one:  python tools/gen_presets.py --robot walter_sr --filepath DIR

Usage (in the build container, where /root/reference exists):
    python tools/gen_presets.py --reference /root/reference
The outputs are committed so that nothing at run time needs the reference.
"""
import argparse
import json
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "operational-space-control_b200")

# MuJoCo model dimensions (W/WW nq=15,nv=14,nu=8; G nq=19,nv=18,nu=12)
MODEL_DIMS = {
    "walter_sr": dict(nq=15, nv=14, nu=8),
    "walter_sr_wheels": dict(nq=15, nv=14, nu=8),
    "unitree_go2": dict(nq=19, nv=18, nu=12),
}
# weight-key prefix of site i, in site order (= noncontact list then contact list):
#   walter_sr/autogen/autogen.py:167-168,187-329 ; unitree_go2/autogen/autogen.py:160-161,178-219
SITE_KEYS = {
    "walter_sr": ["torso", "tls", "trs", "hls", "hrs", "tlh", "trh", "hlh", "hrh",
                  "tlf", "tlr", "trf", "trr", "hlf", "hlr", "hrf", "hrr"],
    "unitree_go2": ["base", "fr", "fl", "hr", "hl"],
}
SITE_KEYS["walter_sr_wheels"] = SITE_KEYS["walter_sr"]
# torque bounds: walter_sr/operational_space_controller.h:309-320 (same in wheels),
#                unitree_go2/operational_space_controller.h:285-296
U_BOUNDS = {
    "walter_sr": ([-1000.0] * 8, [1000.0] * 8),
    "walter_sr_wheels": ([-1000.0] * 8, [1000.0] * 8),
    "unitree_go2": ([-23.7, -23.7, -45.3] * 4, [23.7, 23.7, 45.3] * 4),
}
FZ_MAX = 1e4  # big_number, walter_sr/operational_space_controller.h:282

# preset name -> (robot family, yaml path relative to the reference root)
PRESETS = {
    "walter_sr": ("walter_sr", "config/walter_sr/walter_sr_config.yaml"),
    "walter_sr_wheels": ("walter_sr_wheels", "config/walter_sr_wheels/walter_sr_wheels_config.yaml"),
    "unitree_go2": ("unitree_go2", "config/unitree_go2/unitree_go2_config.yaml"),
    "walter_sr_true_tumbling_mjjoint": ("walter_sr", "config/walter_sr/true_tumbling_mjjoint.yaml"),
    "walter_sr_stairclimbing_slightlybetter": ("walter_sr", "config/walter_sr/stairclimbing_slightlybetter.yaml"),
    "walter_sr_slowtumbling": ("walter_sr", "config/walter_sr/slowtumbling.yaml"),
}


def preset_from_yaml(robot, path):
    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    keys = SITE_KEYS[robot]
    sites = list(cfg["noncontact_site_list"]) + list(cfg["contact_site_list"])
    assert len(sites) == len(keys) == len(cfg["body_list"]), (robot, path)
    w = cfg["weights_config"]
    dims = MODEL_DIMS[robot]
    lb, ub = U_BOUNDS[robot]
    return {
        "robot": robot,
        "source_yaml": os.path.relpath(path, start=os.path.commonpath([path, "/root/reference"]))
        if path.startswith("/root/reference") else os.path.basename(path),
        **dims,
        "ns": len(sites),
        "nc": len(cfg["contact_site_list"]),
        "body_list": list(cfg["body_list"]),
        "noncontact_site_list": list(cfg["noncontact_site_list"]),
        "contact_site_list": list(cfg["contact_site_list"]),
        "mu": float(cfg["friction_coefficient"]),
        "w_trans": [float(w[f"{k}_translational_tracking"]) for k in keys],
        "w_rot": [float(w[f"{k}_rotational_tracking"]) for k in keys],
        "w_torque": float(w["torque"]),
        "w_reg": float(w["regularization"]),
        "u_lb": lb,
        "u_ub": ub,
        "fz_max": FZ_MAX,
    }


def defines_header(p):
    """Same constant names/values as the reference template (autogen.py:470-514)."""
    nv, nu, nc = p["nv"], p["nu"], p["nc"]
    nz = 3 * nc
    n = nv + nu + nz
    q = lambda names: ", ".join(f'"{s}"sv' for s in names)
    sites = p["noncontact_site_list"] + p["contact_site_list"]
    return f"""#pragma once
// GENERATED by tools/gen_presets.py from {p['source_yaml']} -- do not edit.
// Same names and values as the reference's autogen_defines.h
// (operational-space-control/{p['robot']}/autogen/autogen.py:469-513).
#include <array>
#include <string_view>

using namespace std::string_view_literals;

namespace operational_space_controller::constants {{
    namespace model {{
        // Mujoco Model Constants:
        constexpr int nq_size  = {p['nq']};
        constexpr int nv_size = {nv};
        constexpr int nu_size  = {nu};
        constexpr int body_ids_size = {len(p['body_list'])};
        constexpr int site_ids_size = {len(sites)};
        constexpr int noncontact_site_ids_size = {len(p['noncontact_site_list'])};
        constexpr int contact_site_ids_size = {nc};
        constexpr std::array body_list = {{{q(p['body_list'])}}};
        constexpr std::array site_list = {{{q(sites)}}};
        constexpr std::array noncontact_site_list = {{{q(p['noncontact_site_list'])}}};
        constexpr std::array contact_site_list = {{{q(p['contact_site_list'])}}};
    }}
    namespace optimization {{
        // Optimization Constants:
        constexpr int dv_size = {nv};
        constexpr int u_size = {nu};
        constexpr int z_size = {nz};
        constexpr int design_vector_size = {n};
        constexpr int dv_idx = {nv};
        constexpr int u_idx = {nv + nu};
        constexpr int z_idx = {n};
        constexpr int beq_sz = {nv};
        constexpr int Aeq_sz = {nv * n};
        constexpr int Aeq_rows = {nv};
        constexpr int Aeq_cols = {n};
        constexpr int bineq_sz = {4 * nc};
        constexpr int Aineq_sz = {4 * nc * n};
        constexpr int Aineq_rows = {4 * nc};
        constexpr int Aineq_cols = {n};
        constexpr int H_sz = {n * n};
        constexpr int H_rows = {n};
        constexpr int H_cols = {n};
        constexpr int f_sz = {n};
    }}
    // B200 build only: the YAML's objective weights and friction coefficient, which the
    // reference bakes into the CasADi-generated code instead of exposing as constants.
    namespace weights {{
        constexpr double friction_coefficient = {p['mu']!r};
        constexpr std::array<double, {len(sites)}> translational = {{{", ".join(repr(v) for v in p['w_trans'])}}};
        constexpr std::array<double, {len(sites)}> rotational = {{{", ".join(repr(v) for v in p['w_rot'])}}};
        constexpr double torque = {p['w_torque']!r};
        constexpr double regularization = {p['w_reg']!r};
    }}
}}
"""


FUNCS = ("beq", "Aeq", "bineq", "Aineq", "H", "f")


def functions_header(p):
    decl = []
    for fn in FUNCS:
        nin = 1 if fn in ("bineq", "Aineq") else 4
        decl.append(f"""int {fn}(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem);
int {fn}_alloc_mem(void);
int {fn}_init_mem(int mem);
void {fn}_free_mem(int mem);
int {fn}_checkout(void);
void {fn}_release(int mem);
void {fn}_incref(void);
void {fn}_decref(void);
casadi_int {fn}_n_in(void);
casadi_int {fn}_n_out(void);
int {fn}_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w);
#define {fn}_SZ_ARG {nin}
#define {fn}_SZ_RES 1
#define {fn}_SZ_IW 0
#define {fn}_SZ_W 0
""")
    return f"""/* GENERATED by tools/gen_presets.py from {p['source_yaml']} -- do not edit.
 * Same C interface as the CasADi-generated autogen_functions.h of the reference
 * (operational-space-control/{p['robot']}/autogen/autogen.py:440-467), closed-form bodies. */
#ifdef __cplusplus
extern "C" {{
#endif

#ifndef casadi_real
#define casadi_real double
#endif

#ifndef casadi_int
#define casadi_int long long int
#endif

{chr(10).join(decl)}
#ifdef __cplusplus
}} /* extern "C" */
#endif
"""


def functions_source(p):
    nv, nu, nc, ns = p["nv"], p["nu"], p["nc"], p["ns"]
    nz, s6 = 3 * nc, 6 * p["ns"]
    n = nv + nu + nz
    w_row = [p["w_trans"][i // 3] for i in range(3 * ns)] + [p["w_rot"][i // 3] for i in range(3 * ns)]
    boiler = []
    for fn in FUNCS:
        nin = 1 if fn in ("bineq", "Aineq") else 4
        boiler.append(f"""int {fn}_alloc_mem(void) {{ return 0; }}
int {fn}_init_mem(int mem) {{ (void)mem; return 0; }}
void {fn}_free_mem(int mem) {{ (void)mem; }}
int {fn}_checkout(void) {{ return 0; }}
void {fn}_release(int mem) {{ (void)mem; }}
void {fn}_incref(void) {{}}
void {fn}_decref(void) {{}}
casadi_int {fn}_n_in(void) {{ return {nin}; }}
casadi_int {fn}_n_out(void) {{ return 1; }}
int {fn}_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w) {{
  if (sz_arg) *sz_arg = {nin};
  if (sz_res) *sz_res = 1;
  if (sz_iw) *sz_iw = 0;
  if (sz_w) *sz_w = 0;
  return 0;
}}
""")
    return f"""/* GENERATED by tools/gen_presets.py from {p['source_yaml']} -- do not edit.
 * Closed forms of the six CasADi functions of the reference
 * (operational-space-control/{p['robot']}/autogen/autogen.py:54-133 constraints, :135-345
 * objective, :381-426 the functions), with CasADi's conventions: inputs and outputs dense,
 * COLUMN-major; a NULL input is read as zeros; a NULL output is skipped.
 *   eq(q)   = M dv + C - B u - Jc z,  B = [0; I_nu]        Aeq = d eq/dq,  beq = -eq(q)
 *   ineq(q) = friction pyramid (+-1, +-1, -mu) per contact  Aineq = d ineq/dq, bineq = -ineq(q)
 *   obj(q)  = sum_k w_k (J dv + bias - t)_k^2 + w_torque |u|^2 + w_reg |q|^2
 *             H = Hessian, f = gradient AT q   (the reference evaluates at q = 0)           */
#include "autogen_functions.h"

#define NV {nv}
#define NU {nu}
#define NC {nc}
#define NS {ns}
#define NZ {nz}
#define N {n}
#define S6 {s6}

static const casadi_real kMu = {p['mu']!r};
static const casadi_real kWTorque = {p['w_torque']!r};
static const casadi_real kWReg = {p['w_reg']!r};
/* weight of row k of ddx = J dv + bias: 3 NS translational rows, then 3 NS rotational rows */
static const casadi_real kWRow[S6] = {{{", ".join(repr(float(v)) for v in w_row)}}};

static casadi_real in(const casadi_real* a, int i) {{ return a ? a[i] : 0.0; }}

/* arg: q[N], M[NV x NV], C[NV], Jc[NV x NZ] -- res: -(M dv + C - B u - Jc z) */
int beq(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {{
  (void)iw; (void)w; (void)mem;
  if (!res[0]) return 0;
  for (int i = 0; i < NV; ++i) {{
    casadi_real e = in(arg[2], i);
    for (int j = 0; j < NV; ++j) e += in(arg[1], i + NV * j) * in(arg[0], j);
    if (i >= NV - NU) e -= in(arg[0], NV + i - (NV - NU));
    for (int k = 0; k < NZ; ++k) e -= in(arg[3], i + NV * k) * in(arg[0], NV + NU + k);
    res[0][i] = -e;
  }}
  return 0;
}}

/* res: [M, -B, -Jc], NV x N column-major */
int Aeq(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {{
  (void)iw; (void)w; (void)mem;
  if (!res[0]) return 0;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < NV; ++i) {{
      casadi_real v;
      if (j < NV) v = in(arg[1], i + NV * j);
      else if (j < NV + NU) v = (i == (NV - NU) + (j - NV)) ? -1.0 : 0.0;
      else v = -in(arg[3], i + NV * (j - NV - NU));
      res[0][i + NV * j] = v;
    }}
  return 0;
}}

static casadi_real pyramid(int r, int k) {{ /* row r (0..3) of a contact, component k */
  if (k == 0) return (r & 1) ? -1.0 : 1.0;
  if (k == 1) return (r & 2) ? -1.0 : 1.0;
  return -kMu;
}}

/* arg: q -- res: -ineq(q), 4 NC */
int bineq(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {{
  (void)iw; (void)w; (void)mem;
  if (!res[0]) return 0;
  for (int c = 0; c < NC; ++c)
    for (int r = 0; r < 4; ++r) {{
      casadi_real e = 0.0;
      for (int k = 0; k < 3; ++k) e += pyramid(r, k) * in(arg[0], NV + NU + 3 * c + k);
      res[0][4 * c + r] = -e;
    }}
  return 0;
}}

/* res: 4 NC x N column-major */
int Aineq(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {{
  (void)arg; (void)iw; (void)w; (void)mem;
  if (!res[0]) return 0;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < 4 * NC; ++i) {{
      const int c = i / 4, r = i % 4, k = j - (NV + NU + 3 * c);
      res[0][i + 4 * NC * j] = (k >= 0 && k < 3) ? pyramid(r, k) : 0.0;
    }}
  return 0;
}}

/* target of row k of ddx: desired_task_ddx is NS x 6 column-major, translational rows read
 * columns 0-2, rotational rows columns 3-5 (autogen.py:163,173-177) */
static casadi_real target(const casadi_real* t, int k) {{
  const int kr = k < 3 * NS ? k : k - 3 * NS;
  const int site = kr / 3, col = kr % 3 + (k < 3 * NS ? 0 : 3);
  return in(t, site + NS * col);
}}

/* arg: q, desired_task_ddx[NS x 6], J_task[S6 x NV], task_bias[S6] -- res: N x N column-major */
int H(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {{
  (void)iw; (void)w; (void)mem;
  if (!res[0]) return 0;
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) {{
      casadi_real v = 0.0;
      if (i < NV && j < NV)
        for (int k = 0; k < S6; ++k)
          v += 2.0 * kWRow[k] * in(arg[2], k + S6 * i) * in(arg[2], k + S6 * j);
      if (i == j) v += 2.0 * kWReg + ((i >= NV && i < NV + NU) ? 2.0 * kWTorque : 0.0);
      res[0][i + N * j] = v;
    }}
  return 0;
}}

/* res: gradient of the objective at q */
int f(const casadi_real** arg, casadi_real** res, casadi_int* iw, casadi_real* w, int mem) {{
  (void)iw; (void)w; (void)mem;
  if (!res[0]) return 0;
  casadi_real r[S6];
  for (int k = 0; k < S6; ++k) {{
    casadi_real e = in(arg[3], k) - target(arg[1], k);
    for (int j = 0; j < NV; ++j) e += in(arg[2], k + S6 * j) * in(arg[0], j);
    r[k] = kWRow[k] * e;
  }}
  for (int i = 0; i < N; ++i) {{
    casadi_real g = 2.0 * kWReg * in(arg[0], i);
    if (i < NV)
      for (int k = 0; k < S6; ++k) g += 2.0 * in(arg[2], k + S6 * i) * r[k];
    else if (i < NV + NU)
      g += 2.0 * kWTorque * in(arg[0], i);
    res[0][i] = g;
  }}
  return 0;
}}

{chr(10).join(boiler)}"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--robot", default=None,
                    help="with --filepath: emit the genrule outputs of this preset only")
    ap.add_argument("--filepath", default=None,
                    help="output directory of autogen_functions.cc/.h + autogen_defines.h "
                         "(the reference genrule's --filepath=$(RULEDIR))")
    args = ap.parse_args()
    if args.filepath:
        name = args.robot or "walter_sr"
        robot, rel = PRESETS[name]
        p = preset_from_yaml(robot, os.path.join(args.reference, rel))
        p["source_yaml"] = rel
        os.makedirs(args.filepath, exist_ok=True)
        for fn, text in (("autogen_defines.h", defines_header(p)),
                         ("autogen_functions.h", functions_header(p)),
                         ("autogen_functions.cc", functions_source(p))):
            with open(os.path.join(args.filepath, fn), "w") as fh:
                fh.write(text)
        print("wrote genrule outputs of", name, "to", args.filepath)
        return
    os.makedirs(os.path.join(PKG, "presets"), exist_ok=True)
    for name, (robot, rel) in PRESETS.items():
        p = preset_from_yaml(robot, os.path.join(args.reference, rel))
        p["source_yaml"] = rel
        with open(os.path.join(PKG, "presets", f"{name}.json"), "w") as fh:
            json.dump(p, fh, indent=1)
            fh.write("\n")
        if name == robot:  # the wired configs get the C++ header too
            d = os.path.join(PKG, robot, "autogen")
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "autogen_defines.h"), "w") as fh:
                fh.write(defines_header(p))
            with open(os.path.join(d, "autogen_functions.h"), "w") as fh:
                fh.write(functions_header(p))
            with open(os.path.join(d, "autogen_functions.cc"), "w") as fh:
                fh.write(functions_source(p))
        print("wrote", name)


if __name__ == "__main__":
    main()
