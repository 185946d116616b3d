// examples/standing_batched.cc -- N Walter Sr robots standing, driven like the reference's
// examples/walter_sr_standing.cc:89-167 drives one: fill State / OSCData / TaskspaceTargets,
// hand them to the controller, read the torque command.  The records come from a synthetic
// generator here (MuJoCo and the MJCF model are not part of this repository); with MuJoCo on the
// include path `OperationalSpaceController` computes OSCData itself, exactly as the reference.
//
//   g++ -std=c++20 -O2 -I include examples/standing_batched.cc -o standing_batched \
//       -L operational-space-control_b200 -losc_b200 -Wl,-rpath,$PWD/operational-space-control_b200
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "operational-space-control/walter_sr/operational_space_controller.h"

using namespace operational_space_controller::constants;

int main(int argc, char** argv) {
  const int n_envs = argc > 1 ? std::atoi(argv[1]) : 1024;
  BatchedOperationalSpaceController batch(n_envs);
  if (!batch.ok()) {
    std::printf("no usable B200: %s\n", batch.last_error().c_str());
    return 1;
  }
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1u << 24) - 0.5; };
  for (int e = 0; e < n_envs; ++e) {
    OSCData d;
    TaskspaceTargets t = TaskspaceTargets::Zero();
    State st;
    st.contact_mask.setConstant(1.0);  // standing: every wheel on the ground (:107)
    // a consistent synthetic record: J random with the floating-base structure,
    // M = eps I + sum_i m_i Jp_i' Jp_i (SPD), C = gravity pulled through J
    for (int r = 0; r < optimization::s_size; ++r)
      for (int c = 0; c < model::nv_size; ++c) d.taskspace_jacobian(r, c) = 0.3 * rnd();
    for (int r = 0; r < model::nv_size; ++r)
      for (int c = 0; c < model::nv_size; ++c) d.mass_matrix(r, c) = r == c ? 1e-3 : 0.0;
    for (int i = 0; i < model::site_ids_size; ++i) {
      const double m = 0.05 + 1.45 * (rnd() + 0.5);
      for (int k = 0; k < 3; ++k) {
        const int row = 3 * i + k;
        for (int a = 0; a < model::nv_size; ++a)
          for (int b = 0; b < model::nv_size; ++b)
            d.mass_matrix(a, b) += m * d.taskspace_jacobian(row, a) * d.taskspace_jacobian(row, b);
      }
      for (int a = 0; a < model::nv_size; ++a)
        d.coriolis_matrix(a) += m * 9.81 * d.taskspace_jacobian(3 * i + 2, a);
    }
    for (int r = 0; r < optimization::s_size; ++r) d.taskspace_bias(r) = rnd();
    batch.set_environment(e, d, t, st);
  }
  if (!batch.initialize_optimization().ok()) return 1;
  for (int step = 0; step < 10; ++step)
    if (!batch.step().ok()) return 1;
  const auto tau = batch.get_torque_command(0);
  std::printf("%d robots, 10 control steps; torque command of robot 0:", n_envs);
  for (int i = 0; i < model::nu_size; ++i) std::printf(" %.4f", tau(i));
  std::printf("\n");
  for (int i = 0; i < model::nu_size; ++i)
    if (!std::isfinite(tau(i))) return 1;
  return 0;
}
