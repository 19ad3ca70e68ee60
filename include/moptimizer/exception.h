// Mirrors include/moptimizer/exception.h:7-19 of the reference.
#pragma once

#include <exception>
#include <string>

namespace moptimizer {

class Exception : public std::exception {
 public:
  explicit Exception(const char* message) : msg_(message) {}
  explicit Exception(const std::string& message) : msg_(message) {}
  ~Exception() noexcept override = default;
  const char* what() const noexcept override { return msg_.c_str(); }

 protected:
  std::string msg_;
};

}  // namespace moptimizer
