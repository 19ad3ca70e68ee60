// moptimizer::Exception — what the reference throws for API misuse (include/moptimizer/exception.h:7-19: a
// std::exception constructible from a C string or a std::string whose what() returns that text).  Here it also
// carries the C-ABI status of the call that failed, so a caller can tell a bad argument from a CUDA failure.
#pragma once

#include <exception>
#include <string>
#include <utility>

namespace moptimizer {

class Exception : public std::exception {
  std::string text_;  // returned by what()
  int code_ = 0;      // mopt_status of the C-ABI call behind this exception; 0: raised by the C++ layer itself

 public:
  explicit Exception(std::string message, int status = 0) : text_(std::move(message)), code_(status) {}
  explicit Exception(const char* message, int status = 0) : Exception(std::string(message ? message : ""), status) {}

  const char* what() const noexcept override { return text_.c_str(); }
  int status() const noexcept { return code_; }
};

}  // namespace moptimizer
