// tools/reference_dump/osc_dump.h -- header-only recorder for the REFERENCE controller
// (vannem95/operational-space-control), to be built by someone who has its toolchain (Bazel,
// MuJoCo, OSQP): it writes what crosses the boundary of the hot path at every control step,
//   in : OSCData (mass_matrix, coriolis_matrix, taskspace_jacobian, taskspace_bias),
//        taskspace_targets, state.contact_mask                     (:515-555)
//   out: solution, dual_solution, exit_code, iterations             (:589-594)
// to a flat binary file that tools/reference_dump/dump_to_npz.py turns into
// tests/golden/reference_<name>.npz, which tests/test_reference_vectors.py replays through
// the oracle (CPU suite) and the CUDA path (-m gpu).  That is the pin this repository cannot
// make itself: the reference's own outputs on its own inputs.
//
// Only the standard library is used; the hook takes raw pointers, so the three lines added
// to the reference header (README.md in this directory) do not depend on Eigen versions.
//
// File: "OSCDUMP1", int32 {nv, nu, nc, ns, n, m}, then records
//   int32 kind (0 = set_up_optimization / Init, 1 = control step)
//   double M[nv*nv], C[nv], J[6 ns nv], bias[6 ns], targets[ns*6], mask[nc]   (row-major)
//   kind 1 only: double solution[n], dual[m]; int32 exit_code, iterations
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace osc_dump {

struct Shape { int32_t nv, nu, nc, ns, n, m; };

class Recorder {
 public:
  // path from the environment (OSC_DUMP_FILE) so that the examples need no new flag;
  // OSC_DUMP_MAX_STEPS bounds the file (default 2000 control steps)
  static Recorder& instance() { static Recorder r; return r; }
  bool active() const { return f_ != nullptr; }

  void open(const Shape& s) {
    std::lock_guard<std::mutex> lock(mu_);
    if (f_ || tried_) return;
    tried_ = true;
    const char* path = std::getenv("OSC_DUMP_FILE");
    if (!path) return;
    if (const char* mx = std::getenv("OSC_DUMP_MAX_STEPS")) max_steps_ = std::atol(mx);
    f_ = std::fopen(path, "wb");
    if (!f_) return;
    shape_ = s;
    std::fwrite("OSCDUMP1", 1, 8, f_);
    std::fwrite(&shape_, sizeof(shape_), 1, f_);
  }

  // kind 0: call at the end of set_up_optimization() (after solver.Init)
  void record_init(const double* M, const double* C, const double* J, const double* bias,
                   const double* targets, const double* mask) {
    std::lock_guard<std::mutex> lock(mu_);
    if (!f_) return;
    inputs(0, M, C, J, bias, targets, mask);
    std::fflush(f_);
  }
  // kind 1: call in control_loop() right after solve_optimization()
  void record_step(const double* M, const double* C, const double* J, const double* bias,
                   const double* targets, const double* mask, const double* solution,
                   const double* dual, int exit_code, int iterations) {
    std::lock_guard<std::mutex> lock(mu_);
    if (!f_ || steps_ >= max_steps_) return;
    inputs(1, M, C, J, bias, targets, mask);
    std::fwrite(solution, sizeof(double), (size_t)shape_.n, f_);
    std::fwrite(dual, sizeof(double), (size_t)shape_.m, f_);
    const int32_t tail[2] = {exit_code, iterations};
    std::fwrite(tail, sizeof(int32_t), 2, f_);
    if ((++steps_ & 63) == 0) std::fflush(f_);
  }
  ~Recorder() { if (f_) std::fclose(f_); }

 private:
  void inputs(int32_t kind, const double* M, const double* C, const double* J,
              const double* bias, const double* targets, const double* mask) {
    const size_t nv = shape_.nv, s = 6 * (size_t)shape_.ns;
    std::fwrite(&kind, sizeof(kind), 1, f_);
    std::fwrite(M, sizeof(double), nv * nv, f_);
    std::fwrite(C, sizeof(double), nv, f_);
    std::fwrite(J, sizeof(double), s * nv, f_);
    std::fwrite(bias, sizeof(double), s, f_);
    std::fwrite(targets, sizeof(double), (size_t)shape_.ns * 6, f_);
    std::fwrite(mask, sizeof(double), (size_t)shape_.nc, f_);
  }
  std::FILE* f_ = nullptr;
  bool tried_ = false;
  Shape shape_{};
  long steps_ = 0, max_steps_ = 2000;
  std::mutex mu_;
};

}  // namespace osc_dump
