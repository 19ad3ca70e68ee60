// moptimizer::Exception — what the reference throws for API misuse (include/moptimizer/exception.h:7-19).  Here it
// also carries the C-ABI status of the call that failed, so a caller can tell a bad argument from a CUDA failure.
#pragma once

#include <exception>
#include <string>
#include <utility>

namespace moptimizer {

class Exception : public std::exception {
 public:
  explicit Exception(std::string message, int status = 0) : msg_(std::move(message)), status_(status) {}
  explicit Exception(const char* message, int status = 0) : msg_(message ? message : ""), status_(status) {}
  const char* what() const noexcept override { return msg_.c_str(); }
  /// mopt_status of the C-ABI call behind this exception, 0 when it was raised by the C++ layer itself.
  int status() const noexcept { return status_; }

 protected:
  std::string msg_;
  int status_;
};

}  // namespace moptimizer
