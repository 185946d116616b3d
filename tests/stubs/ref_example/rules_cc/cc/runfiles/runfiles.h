// tests/stubs/ref_example/rules_cc/cc/runfiles/runfiles.h -- Bazel runfiles look-up, as the examples call it.
#pragma once
#include <memory>
#include <string>
#ifndef BAZEL_CURRENT_REPOSITORY
#define BAZEL_CURRENT_REPOSITORY ""
#endif
namespace rules_cc::cc::runfiles {
class Runfiles {
 public:
  static Runfiles* Create(const std::string&, const std::string&, std::string*) { return new Runfiles(); }
  std::string Rlocation(const std::string& path) const { return path; }
};
}  // namespace rules_cc::cc::runfiles
