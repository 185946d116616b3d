// tests/stubs/ref_example/absl/log/absl_check.h -- ABSL_CHECK(cond) << message;
#pragma once
#include <cstdlib>
#include <iostream>
namespace osc_b200_stub {
struct CheckSink {
  bool failed;
  template <class T> CheckSink& operator<<(const T& v) { if (failed) std::cerr << v; return *this; }
  ~CheckSink() { if (failed) { std::cerr << std::endl; std::abort(); } }
};
}  // namespace osc_b200_stub
#define ABSL_CHECK(cond) ::osc_b200_stub::CheckSink{!(cond)}
