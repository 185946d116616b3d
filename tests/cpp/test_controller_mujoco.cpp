// tests/cpp/test_controller_mujoco.cpp -- the OSC_B200_HAVE_MUJOCO branch of the drop-in
// classes (update_mj_data / update_osc_data, reference
// walter_sr/operational_space_controller.h:394-513), compiled against tests/stubs/mujoco/mujoco.h
// and run on tests/stubs/fake_mujoco.cc: the controller is driven exactly like
// examples/walter_sr_standing.cc:89-117 (constructor with a model path, initialize,
// initialize_optimization), with NO OSCData injected -- M, C, J, bias reach the GPU through
// mj_fullM / qfrc_bias / mj_jac / mj_jacDot.
#include <cstdio>
#include <cstdlib>
#include <vector>

#if defined(ROBOT_GO2)
#include "operational-space-control/unitree_go2/operational_space_controller.h"
#elif defined(ROBOT_WW)
#include "operational-space-control/walter_sr_wheels/operational_space_controller.h"
#else
#include "operational-space-control/walter_sr/operational_space_controller.h"
#endif

#if !OSC_B200_HAVE_MUJOCO
#error "compile with -I tests/stubs so that mujoco/mujoco.h is found"
#endif

extern "C" const double* fake_mj_last_qpos(int* n);
extern "C" const double* fake_mj_last_qvel(int* n);
extern "C" int fake_mj_forward_calls();

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  // argv[1]: model blob for the fake MuJoCo; argv[2]: targets + mask
  TaskspaceTargets targets;
  State state;
  {
    FILE* f = fopen(argv[2], "rb");
    if (!f) return 2;
    bool ok = fread(targets.data(), 8, model::site_ids_size * 6, f) == (size_t)model::site_ids_size * 6 &&
              fread(state.contact_mask.data(), 8, model::contact_site_ids_size, f) ==
                  (size_t)model::contact_site_ids_size;
    fclose(f);
    if (!ok) return 2;
  }
  for (int i = 0; i < model::nu_size; ++i) {
    state.motor_position(i) = 0.1 * (i + 1);
    state.motor_velocity(i) = 0.0;
  }
  state.body_rotation(0) = 0.5; state.body_rotation(1) = -0.5;
  state.body_rotation(2) = 0.5; state.body_rotation(3) = 0.5;
  state.linear_body_velocity(0) = 1.0;  // the fake's Jdot is scripted for qvel = e_0

  OperationalSpaceController bad("/nonexistent/model.xml");
  if (bad.initialize(state).ok()) { std::puts("expected: Failed to load Mujoco Model"); return 1; }

  OperationalSpaceController controller(argv[1], 2000);
  absl::Status result = controller.initialize(state);
  if (!result.ok()) { std::printf("initialize: %s\n", std::string(result.message()).c_str()); return 1; }
  controller.update_taskspace_targets(targets);
  result = controller.initialize_optimization();
  if (!result.ok()) { std::printf("initialize_optimization: %s\n", std::string(result.message()).c_str()); return 1; }
  result = controller.step_once();
  if (!result.ok()) { std::printf("step: %s\n", std::string(result.message()).c_str()); return 1; }
  auto torque = controller.get_torque_command();
  std::printf("TORQUE");
  for (int i = 0; i < model::nu_size; ++i) std::printf(" %.17g", torque(i));
  std::printf("\n");
  // what update_mj_data handed to MuJoCo: qpos = [0 0 0, quat, motor_position] (:402-404)
  int nq = 0, nv = 0;
  const double* qpos = fake_mj_last_qpos(&nq);
  const double* qvel = fake_mj_last_qvel(&nv);
  bool ok = nq == model::nq_size && nv == model::nv_size;
  for (int i = 0; ok && i < 3; ++i) ok = qpos[i] == 0.0;
  for (int i = 0; ok && i < 4; ++i) ok = qpos[3 + i] == state.body_rotation(i);
  for (int i = 0; ok && i < model::nu_size; ++i) ok = qpos[7 + i] == state.motor_position(i);
  ok = ok && qvel[0] == 1.0;
  for (int i = 1; ok && i < model::nv_size; ++i) ok = qvel[i] == 0.0;
  std::printf("QPOS_OK %d\nFORWARD_CALLS %d\n", ok ? 1 : 0, fake_mj_forward_calls());
  result = controller.clean_up();
  return result.ok() ? 0 : 1;
}
