// compat.h -- the third-party types that appear in the reference controller's public API
// (absl::Status, osqp::OsqpSettings, fixed-size Eigen matrices).  The real headers are used
// when they are on the include path; otherwise small layout-compatible stand-ins are
// defined so that the controller classes build in an image that has none of them
// (this one: SURVEY.md Appendix C).
#pragma once

#include <array>
#include <cstddef>
#include <initializer_list>
#include <string>
#include <string_view>

// ---------------------------------------------------------------- absl::Status
#if __has_include("absl/status/status.h")
#include "absl/status/status.h"
#else
namespace absl {
enum class StatusCode { kOk = 0, kInvalidArgument = 3, kFailedPrecondition = 9, kInternal = 13 };
class Status {
 public:
  Status() = default;
  Status(StatusCode c, std::string_view m) : code_(c), msg_(m) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  std::string_view message() const { return msg_; }
  void Update(const Status& s) {
    if (ok()) *this = s;
  }
 private:
  StatusCode code_ = StatusCode::kOk;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
inline Status InternalError(std::string_view m) { return Status(StatusCode::kInternal, m); }
inline Status FailedPreconditionError(std::string_view m) { return Status(StatusCode::kFailedPrecondition, m); }
inline Status InvalidArgumentError(std::string_view m) { return Status(StatusCode::kInvalidArgument, m); }
}  // namespace absl
#endif

// ---------------------------------------------------------------- osqp::OsqpSettings
#if __has_include("osqp++.h")
#include "osqp++.h"
#else
namespace osqp {
// Same field names and OSQP 0.6.3 defaults as osqp-cpp's OsqpSettings.
struct OsqpSettings {
  double rho = 0.1;
  double sigma = 1e-6;
  long long scaling = 10;
  bool adaptive_rho = true;
  long long adaptive_rho_interval = 0;
  double adaptive_rho_tolerance = 5.0;
  double adaptive_rho_fraction = 0.4;
  long long max_iter = 4000;
  double eps_abs = 1e-3;
  double eps_rel = 1e-3;
  double eps_prim_inf = 1e-4;
  double eps_dual_inf = 1e-4;
  double alpha = 1.6;
  double delta = 1e-6;
  bool polish = false;
  long long polish_refine_iter = 3;
  bool verbose = true;
  bool scaled_termination = false;
  long long check_termination = 25;
  bool warm_start = true;
  double time_limit = 0.0;
};
}  // namespace osqp
#endif

// ---------------------------------------------------------------- fixed-size matrices
#if __has_include("Eigen/Dense")
#include "Eigen/Dense"
#define OSC_B200_HAVE_EIGEN 1
#else
#define OSC_B200_HAVE_EIGEN 0
namespace osc_b200 {
// Contiguous fixed-size matrix of doubles with the subset of Eigen's interface the
// controller API needs.  Layout-compatible with Eigen::Matrix<double,R,C,Order>.
template <int R, int C, bool RowMajorOrder>
class FixedMatrix {
 public:
  FixedMatrix() { v_.fill(0.0); }
  FixedMatrix(std::initializer_list<double> init) {
    v_.fill(0.0);
    std::size_t k = 0;
    for (double d : init) if (k < v_.size()) v_[k++] = d;
  }
  static FixedMatrix Zero() { return FixedMatrix(); }
  static FixedMatrix Constant(double c) { FixedMatrix m; m.v_.fill(c); return m; }
  static FixedMatrix Ones() { return Constant(1.0); }
  void setZero() { v_.fill(0.0); }
  void setConstant(double c) { v_.fill(c); }
  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }
  static constexpr int size() { return R * C; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(int i, int j) { return v_[RowMajorOrder ? i * C + j : j * R + i]; }
  double operator()(int i, int j) const { return v_[RowMajorOrder ? i * C + j : j * R + i]; }
  double& operator()(int i) { return v_[i]; }
  double operator()(int i) const { return v_[i]; }
  double& operator[](int i) { return v_[i]; }
  double operator[](int i) const { return v_[i]; }
 private:
  std::array<double, static_cast<std::size_t>(R) * C> v_;
};
}  // namespace osc_b200
#endif
