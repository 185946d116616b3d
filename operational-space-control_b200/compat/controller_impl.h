// controller_impl.h -- the reference's OperationalSpaceController on top of the C-ABI.
//
// Public surface = reference walter_sr/operational_space_controller.h:110-240 (ctor,
// initialize, initialize_optimization, initialize_thread, stop_thread, clean_up,
// is_*initialized, update_state, update_taskspace_targets, get_torque_command,
// get_solution), same thread + mutex hand-off (:222-240, :604-647).  The private per-step
// pipeline update_optimization_data -> update_optimization -> solve_optimization ->
// torque slice (:515-631) is replaced by osc_setup / osc_step_host on a one-environment
// handle; `BatchedOperationalSpaceController` below is the N-environment sibling.
//
// MuJoCo (update_mj_data / update_osc_data, :394-513) is upstream of the GPU boundary.
// When mujoco/mujoco.h is on the include path it is called exactly like the reference
// does; otherwise OSCData must be supplied with update_osc_data(const OSCData&).
#pragma once

#include <atomic>
#include <chrono>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/osc_b200.h"
#include "compat.h"

#if __has_include("mujoco/mujoco.h")
#include "mujoco/mujoco.h"
#define OSC_B200_HAVE_MUJOCO 1
#else
#define OSC_B200_HAVE_MUJOCO 0
#endif

namespace osc_b200 {

inline osc_settings to_c_settings(const osqp::OsqpSettings& s) {
  osc_settings c;
  osc_default_settings(&c);
  c.rho = s.rho;
  c.sigma = s.sigma;
  c.alpha = s.alpha;
  c.eps_abs = s.eps_abs;
  c.eps_rel = s.eps_rel;
  c.adaptive_rho_tolerance = s.adaptive_rho_tolerance;
  c.scaling = static_cast<int>(s.scaling);
  c.adaptive_rho = s.adaptive_rho ? 1 : 0;
  c.adaptive_rho_interval = static_cast<int>(s.adaptive_rho_interval);
  c.max_iter = static_cast<int>(s.max_iter);
  c.check_termination = static_cast<int>(s.check_termination);
  c.warm_start = s.warm_start ? 1 : 0;
  c.eps_prim_inf = s.eps_prim_inf;
  c.eps_dual_inf = s.eps_dual_inf;
  return c;
}

// Traits supplies: nv,nu,nc,ns sizes, State, OSCData, TaskspaceTargets, TorqueVector,
// SolutionVector types, and fill_spec(osc_robot_spec&).
template <class Traits>
class Controller {
 public:
  using State = typename Traits::State;
  using OSCData = typename Traits::OSCData;
  using TaskspaceTargets = typename Traits::TaskspaceTargets;
  using TorqueVector = typename Traits::TorqueVector;
  using SolutionVector = typename Traits::SolutionVector;

  Controller(std::filesystem::path xml_path, int control_rate_us = 2000,
             osqp::OsqpSettings osqp_settings = osqp::OsqpSettings())
      : xml_path(std::move(xml_path)), control_rate_us(control_rate_us), settings(osqp_settings) {}
  ~Controller() {
    if (thread_initialized) (void)stop_thread();
    if (handle) osc_destroy(handle);
  }
  Controller(const Controller&) = delete;
  Controller& operator=(const Controller&) = delete;

  absl::Status initialize(State initial_state) {
#if OSC_B200_HAVE_MUJOCO
    char error[1000];
    mj_model = mj_loadXML(xml_path.c_str(), nullptr, error, 1000);
    if (!mj_model) return absl::InternalError("Failed to load Mujoco Model");
    mj_model->opt.timestep = 0.002;
    mj_data = mj_makeData(mj_model);
    absl::Status ids = Traits::resolve_ids(mj_model, site_ids, body_ids);
    if (!ids.ok()) return ids;
#endif
    osc_robot_spec spec;
    Traits::fill_spec(spec);
    const osc_settings cs = to_c_settings(settings);
    if (osc_create(&spec, &cs, 1, device, &handle) != OSC_OK)
      return absl::InternalError(std::string("osc_create: ") + osc_last_error(nullptr));
    state = initial_state;
    initialized = true;
    return absl::OkStatus();
  }

  absl::Status initialize_optimization() {
    if (!initialized)
      return absl::FailedPreconditionError(
          "Operational Space Controller not initialized. Cannot initialize optimization.");
#if OSC_B200_HAVE_MUJOCO
    Traits::update_mj_data(mj_model, mj_data, state, site_ids, points);
    Traits::update_osc_data(mj_model, mj_data, body_ids, points, osc_data);
    have_osc_data = true;
#endif
    if (!have_osc_data)
      return absl::FailedPreconditionError(
          "no OSCData: built without MuJoCo, call update_osc_data(const OSCData&) first");
    absl::Status s = upload();
    if (!s.ok()) return s;
    if (osc_setup(handle, nullptr) != OSC_OK)  // set_up_optimization(): Init
      return absl::InternalError(std::string("osc_setup: ") + osc_last_error(handle));
    optimization_initialized = true;
    return absl::OkStatus();
  }

  absl::Status initialize_thread() {
    if (!initialized || !optimization_initialized)
      return absl::FailedPreconditionError(
          "Operational Space Controller not initialized. Cannot initialize thread.");
    running = true;
    thread = std::thread(&Controller::control_loop, this);
    thread_initialized = true;
    return absl::OkStatus();
  }

  absl::Status stop_thread() {
    if (!initialized || !thread_initialized)
      return absl::FailedPreconditionError(
          "Operational Space Controller not initialized. Cannot stop thread.");
    running = false;
    thread.join();
    thread_initialized = false;
    return absl::OkStatus();
  }

  absl::Status clean_up() {
#if OSC_B200_HAVE_MUJOCO
    if (mj_data) mj_deleteData(mj_data);
    if (mj_model) mj_deleteModel(mj_model);
    mj_data = nullptr;
    mj_model = nullptr;
#endif
    if (handle) {
      osc_destroy(handle);
      handle = nullptr;
    }
    initialized = optimization_initialized = false;
    return absl::OkStatus();
  }

  bool is_initialized() { return initialized; }
  bool is_optimization_initialized() { return optimization_initialized; }
  bool is_thread_initialized() { return thread_initialized; }

  void update_state(const State& new_state) {
    std::lock_guard<std::mutex> lock(mutex);
    state = new_state;
  }
  void update_taskspace_targets(const TaskspaceTargets& new_targets) {
    std::lock_guard<std::mutex> lock(mutex);
    taskspace_targets = new_targets;
  }
  // B200 build only: the MuJoCo-derived record, for hosts that compute it elsewhere.
  void update_osc_data(const OSCData& new_data) {
    std::lock_guard<std::mutex> lock(mutex);
    osc_data = new_data;
    have_osc_data = true;
  }
  TorqueVector get_torque_command() {
    std::lock_guard<std::mutex> lock(mutex);
    return torque_command;
  }
  SolutionVector get_solution() {
    std::lock_guard<std::mutex> lock(mutex);
    return solution;
  }
  // One synchronous control step (what control_loop runs under the mutex).
  absl::Status step_once() {
    std::lock_guard<std::mutex> lock(mutex);
    return step_locked();
  }
  void set_device(int d) { device = d; }
  // B200 build only: control steps completed so far (synchronous + control_loop), so that a
  // test can tell which step a published torque belongs to.
  long long steps_done() {
    std::lock_guard<std::mutex> lock(mutex);
    return steps;
  }

 private:
  absl::Status upload() {
    if (osc_upload(handle, osc_data.mass_matrix.data(), osc_data.coriolis_matrix.data(),
                   osc_data.taskspace_jacobian.data(), osc_data.taskspace_bias.data(),
                   taskspace_targets.data(), state.contact_mask.data(), nullptr) != OSC_OK ||
        osc_sync(handle, nullptr) != OSC_OK)
      return absl::InternalError(std::string("osc_upload: ") + osc_last_error(handle));
    return absl::OkStatus();
  }
  absl::Status step_locked() {
#if OSC_B200_HAVE_MUJOCO
    Traits::update_mj_data(mj_model, mj_data, state, site_ids, points);
    Traits::update_osc_data(mj_model, mj_data, body_ids, points, osc_data);
#endif
    if (osc_step_host(handle, osc_data.mass_matrix.data(), osc_data.coriolis_matrix.data(),
                      osc_data.taskspace_jacobian.data(), osc_data.taskspace_bias.data(),
                      taskspace_targets.data(), state.contact_mask.data(), torque_command.data(),
                      nullptr) != OSC_OK)
      return absl::InternalError(std::string("osc_step_host: ") + osc_last_error(handle));
    if (osc_download(handle, nullptr, solution.data(), nullptr, nullptr, &exit_code, nullptr,
                     nullptr, nullptr, nullptr) != OSC_OK ||
        osc_sync(handle, nullptr) != OSC_OK)
      return absl::InternalError(std::string("osc_download: ") + osc_last_error(handle));
    ++steps;
    return absl::OkStatus();
  }
  void control_loop() {
    using Clock = std::chrono::steady_clock;
    auto next_time = Clock::now();
    while (running) {
      next_time += std::chrono::microseconds(control_rate_us);
      {
        std::lock_guard<std::mutex> lock(mutex);
        std::ignore = step_locked();  // the reference discards the status too (:625)
      }
      auto now = Clock::now();
      if (now < next_time) {
        std::this_thread::sleep_until(next_time);
      } else {
        auto overrun = std::chrono::duration_cast<std::chrono::microseconds>(now - next_time);
        std::cout << "Operational Space Control Loop Execution Time Exceeded Control Rate: "
                  << overrun.count() << "us" << std::endl;
        next_time = now;
      }
    }
  }

  State state;
  TaskspaceTargets taskspace_targets = TaskspaceTargets::Zero();
  TorqueVector torque_command = TorqueVector::Zero();
  SolutionVector solution = SolutionVector::Zero();
  OSCData osc_data;
  bool have_osc_data = false;
  bool initialized = false, optimization_initialized = false, thread_initialized = false;
  std::filesystem::path xml_path;
  int control_rate_us;
  osqp::OsqpSettings settings;
  int exit_code = OSC_UNSOLVED;
  long long steps = 0;
  int device = 0;
  osc_handle* handle = nullptr;
  std::atomic<bool> running{true};
  std::mutex mutex;
  std::thread thread;
#if OSC_B200_HAVE_MUJOCO
  mjModel* mj_model = nullptr;
  mjData* mj_data = nullptr;
  std::vector<int> site_ids, body_ids;
  typename Traits::Points points;
#endif
};

// N environments at once: arrays of the reference's records, one call per control step.
template <class Traits>
class BatchedController {
 public:
  using State = typename Traits::State;
  using OSCData = typename Traits::OSCData;
  using TaskspaceTargets = typename Traits::TaskspaceTargets;
  using TorqueVector = typename Traits::TorqueVector;

  BatchedController(int n_envs, osqp::OsqpSettings osqp_settings = osqp::OsqpSettings(),
                    int device = 0)
      : n_envs(n_envs) {
    osc_robot_spec spec;
    Traits::fill_spec(spec);
    const osc_settings cs = to_c_settings(osqp_settings);
    if (osc_create(&spec, &cs, n_envs, device, &handle) != OSC_OK) {
      error = std::string("osc_create: ") + osc_last_error(nullptr);
      handle = nullptr;
    }
    const int nv = Traits::nv, s = 6 * Traits::ns;
    M.resize((size_t)n_envs * nv * nv);
    C.resize((size_t)n_envs * nv);
    J.resize((size_t)n_envs * s * nv);
    bias.resize((size_t)n_envs * s);
    targets.resize((size_t)n_envs * s);
    mask.resize((size_t)n_envs * Traits::nc);
    torque.resize((size_t)n_envs * Traits::nu);
  }
  ~BatchedController() {
    if (handle) osc_destroy(handle);
  }
  bool ok() const { return handle != nullptr; }
  const std::string& last_error() const { return error; }

  // pack env e's records (AoS, the reference's layouts) into the batch arrays
  void set_environment(int e, const OSCData& d, const TaskspaceTargets& t, const State& st) {
    const int nv = Traits::nv, s = 6 * Traits::ns;
    std::memcpy(&M[(size_t)e * nv * nv], d.mass_matrix.data(), sizeof(double) * nv * nv);
    std::memcpy(&C[(size_t)e * nv], d.coriolis_matrix.data(), sizeof(double) * nv);
    std::memcpy(&J[(size_t)e * s * nv], d.taskspace_jacobian.data(), sizeof(double) * s * nv);
    std::memcpy(&bias[(size_t)e * s], d.taskspace_bias.data(), sizeof(double) * s);
    std::memcpy(&targets[(size_t)e * s], t.data(), sizeof(double) * s);
    std::memcpy(&mask[(size_t)e * Traits::nc], st.contact_mask.data(), sizeof(double) * Traits::nc);
  }
  absl::Status initialize_optimization() {
    if (!handle) return absl::InternalError(error);
    if (osc_upload(handle, M.data(), C.data(), J.data(), bias.data(), targets.data(), mask.data(),
                   nullptr) != OSC_OK || osc_setup(handle, nullptr) != OSC_OK ||
        osc_sync(handle, nullptr) != OSC_OK)
      return absl::InternalError(osc_last_error(handle));
    return absl::OkStatus();
  }
  absl::Status step() {
    if (!handle) return absl::InternalError(error);
    if (osc_step_host(handle, M.data(), C.data(), J.data(), bias.data(), targets.data(),
                      mask.data(), torque.data(), nullptr) != OSC_OK)
      return absl::InternalError(osc_last_error(handle));
    return absl::OkStatus();
  }
  TorqueVector get_torque_command(int e) const {
    TorqueVector t;
    std::memcpy(t.data(), &torque[(size_t)e * Traits::nu], sizeof(double) * Traits::nu);
    return t;
  }
  // ---- roll-outs that keep targets and contact masks on the device (the step before the
  //      path: examples/standing.cc:146-155, examples/walter_sr_true_tumbling_mjjoint.cc:523-558)
  // task-space PD targets of every (environment, site) from DEVICE-resident site states
  absl::Status update_taskspace_targets_pd(const osc_site_state& sites, const double* kp_lin,
                                           const double* kd_lin, const double* kp_ang,
                                           const double* kd_ang, void* stream = nullptr) {
    if (!handle) return absl::InternalError(error);
    if (osc_targets_pd(handle, &sites, kp_lin, kd_lin, kp_ang, kd_ang, stream) != OSC_OK)
      return absl::InternalError(osc_last_error(handle));
    return absl::OkStatus();
  }
  // contact mask from DEVICE-resident MuJoCo contact geom pairs
  absl::Status update_contact_mask_from_contacts(const int* geom_pairs, const int* ncon,
                                                 int max_con, const int* contact_geom_ids,
                                                 const int* site_of_geom = nullptr,
                                                 void* stream = nullptr) {
    if (!handle) return absl::InternalError(error);
    if (osc_contact_mask_from_contacts(handle, geom_pairs, ncon, max_con, contact_geom_ids,
                                       site_of_geom, stream) != OSC_OK)
      return absl::InternalError(osc_last_error(handle));
    return absl::OkStatus();
  }
  // control step on the inputs resident in HBM (uploaded earlier / produced by the two calls
  // above); torques come back to the host
  absl::Status step_resident(void* stream = nullptr) {
    if (!handle) return absl::InternalError(error);
    if (osc_step(handle, stream) != OSC_OK ||
        osc_download(handle, torque.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                     nullptr, stream) != OSC_OK ||
        osc_sync(handle, stream) != OSC_OK)
      return absl::InternalError(osc_last_error(handle));
    return absl::OkStatus();
  }
  osc_handle* c_handle() { return handle; }

 private:
  int n_envs;
  osc_handle* handle = nullptr;
  std::string error;
  std::vector<double> M, C, J, bias, targets, mask, torque;
};

}  // namespace osc_b200
