// tests/cpp/test_controller.cpp -- exercises the drop-in C++ class the way
// examples/walter_sr_standing.cc:89-167 drives the reference controller, with OSCData
// injected (no MuJoCo in this image).  Reads one environment's inputs from a binary file
// written by the pytest wrapper and prints the torques for comparison with the oracle.
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

static bool read(FILE* f, double* p, size_t n) { return fread(p, sizeof(double), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  OSCData data;
  TaskspaceTargets targets;
  State state;
  bool ok = read(f, data.mass_matrix.data(), model::nv_size * model::nv_size) &&
            read(f, data.coriolis_matrix.data(), model::nv_size) &&
            read(f, data.taskspace_jacobian.data(), optimization::s_size * model::nv_size) &&
            read(f, data.taskspace_bias.data(), optimization::s_size) &&
            read(f, targets.data(), model::site_ids_size * 6) &&
            read(f, state.contact_mask.data(), model::contact_site_ids_size);
  fclose(f);
  if (!ok) return 2;

  OperationalSpaceController controller("unused_without_mujoco.xml", 2000);
  absl::Status result;
  result.Update(controller.initialize_optimization());
  if (result.ok()) { std::puts("expected FailedPrecondition before initialize()"); return 1; }
  result = controller.initialize(state);
  if (!result.ok()) { std::printf("initialize: %s\n", std::string(result.message()).c_str()); return 1; }
  controller.update_osc_data(data);
  controller.update_taskspace_targets(targets);
  result = controller.initialize_optimization();
  if (!result.ok()) { std::printf("initialize_optimization: %s\n", std::string(result.message()).c_str()); return 1; }
  // synchronous step (deterministic output for the parity check) ...
  result = controller.step_once();
  if (!result.ok()) { std::printf("step: %s\n", std::string(result.message()).c_str()); return 1; }
  auto torque = controller.get_torque_command();
  auto solution = controller.get_solution();
  std::printf("TORQUE");
  for (int i = 0; i < model::nu_size; ++i) std::printf(" %.17g", torque(i));
  std::printf("\nSLICE_OK %d\n", [&] {
    for (int i = 0; i < model::nu_size; ++i)
      if (solution(optimization::dv_idx + i) != torque(i)) return 0;
    return 1;
  }());
  // ... then the reference's asynchronous life cycle: thread at control_rate_us, mutex hand-off
  result = controller.initialize_thread();
  if (!result.ok() || !controller.is_thread_initialized()) return 1;
  for (int k = 0; k < 20; ++k) {
    controller.update_state(state);
    controller.update_taskspace_targets(targets);
    (void)controller.get_torque_command();
    std::this_thread::sleep_for(std::chrono::milliseconds(2));
  }
  result = controller.stop_thread();
  if (!result.ok()) return 1;
  auto torque2 = controller.get_torque_command();
  // the thread ran steps_done() - 1 warm control steps on the same inputs after the
  // synchronous one: the test replays exactly that many on the oracle
  std::printf("THREAD_STEPS %lld\n", controller.steps_done());
  std::printf("TORQUE_THREAD");
  for (int i = 0; i < model::nu_size; ++i) std::printf(" %.17g", torque2(i));
  std::printf("\n");
  {  // a caller that names the reference's OptimizationData record still compiles (:23-31)
    OptimizationData od;
    std::printf("OPTDATA %d %d\n", (int)od.H.size(), (int)od.Aineq.size());
  }
  result = controller.clean_up();
  if (!result.ok()) return 1;
  // the N-environment sibling: three copies of the same robot, host step then resident step
  {
    BatchedOperationalSpaceController batch(3);
    if (!batch.ok()) { std::printf("batch: %s\n", batch.last_error().c_str()); return 1; }
    for (int e = 0; e < 3; ++e) batch.set_environment(e, data, targets, state);
    if (!batch.initialize_optimization().ok() || !batch.step().ok()) return 1;
    auto tb = batch.get_torque_command(2);
    std::printf("TORQUE_BATCH");
    for (int i = 0; i < model::nu_size; ++i) std::printf(" %.17g", tb(i));
    std::printf("\n");
    if (!batch.step_resident().ok()) return 1;  // second control step on the resident inputs
    auto tr = batch.get_torque_command(0);
    std::printf("TORQUE_RESIDENT");
    for (int i = 0; i < model::nu_size; ++i) std::printf(" %.17g", tr(i));
    std::printf("\n");
  }
  return 0;
}
