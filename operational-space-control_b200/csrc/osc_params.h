// osc_params.h -- C-ABI structs (include/osc_b200.h) -> launch constants (osc::Params).
#pragma once

#include "../../include/osc_b200.h"
#include "osc_core.cuh"
#include "osc_core3.cuh"

namespace osc {

enum class Shape { kNone, kWalter, kGo2 };
using WalterDims = Dims<14, 8, 8, 17>;  // walter_sr, walter_sr_wheels
using Go2Dims = Dims<18, 12, 4, 5>;     // unitree_go2

inline Shape shape_of(const osc_robot_spec& r) {
  if (r.nv == 14 && r.nu == 8 && r.nc == 8 && r.ns == 17) return Shape::kWalter;
  if (r.nv == 18 && r.nu == 12 && r.nc == 4 && r.ns == 5) return Shape::kGo2;
  return Shape::kNone;
}

inline Params make_params(const osc_robot_spec& r, const osc_settings& s) {
  Params p{};
  // rows of ddx = J dv + bias: 3 translational rows per site, then 3 rotational rows per
  // site (autogen.py:163); one weight per (site, block) (autogen.py:187-329)
  for (int i = 0; i < r.ns; ++i)
    for (int k = 0; k < 3; ++k) {
      p.w_row[3 * i + k] = r.w_trans[i];
      p.w_row[3 * r.ns + 3 * i + k] = r.w_rot[i];
    }
  p.w_reg = r.w_reg;
  p.w_torque = r.w_torque;
  p.mu = r.mu;
  p.fz_max = r.fz_max;
  for (int j = 0; j < r.nu; ++j) {
    p.u_lb[j] = r.u_lb[j];
    p.u_ub[j] = r.u_ub[j];
  }
  p.rho0 = s.rho;
  p.sigma = s.sigma;
  p.alpha = s.alpha;
  p.eps_abs = s.eps_abs;
  p.eps_rel = s.eps_rel;
  p.rho_tol = s.adaptive_rho_tolerance;
  p.eps_prim_inf = s.eps_prim_inf;
  p.eps_dual_inf = s.eps_dual_inf;
  p.scaling = s.scaling;
  p.adaptive_rho = s.adaptive_rho;
  p.adaptive_rho_interval = s.adaptive_rho_interval;
  p.max_iter = s.max_iter;
  p.check_termination = s.check_termination;
  p.warm_start = s.warm_start;
  return p;
}

inline void default_settings(osc_settings* s) {
  s->rho = 0.1;
  s->sigma = 1e-6;
  s->alpha = 1.6;
  s->eps_abs = 1e-3;
  s->eps_rel = 1e-3;
  s->adaptive_rho_tolerance = 5.0;
  s->scaling = 10;
  s->adaptive_rho = 1;
  s->adaptive_rho_interval = 0;
  s->max_iter = 4000;
  s->check_termination = 25;
  s->warm_start = 1;
  s->eps_prim_inf = 1e-4;
  s->eps_dual_inf = 1e-4;
}

}  // namespace osc
