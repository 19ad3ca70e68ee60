// Minimal equivalent of include/moptimizer/logger.h:12-65 (duna::Logger), kept only because
// LevenbergMarquadtDynamic::setLogger is part of the optimizer API (levenberg_marquadt_dyn.h:20).
#pragma once

#include <iostream>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <unordered_set>

namespace duna {

class Logger {
 public:
  using LoggerPtr = std::shared_ptr<Logger>;
  enum VERBOSITY_LEVEL { L_ERROR, L_WARN, L_INFO, L_DEBUG };

  Logger(std::ostream& sink, VERBOSITY_LEVEL level = L_ERROR, const std::string& name = "logger")
      : sink_(sink), level_(level), name_(name) {}

  template <typename... Args>
  void log(VERBOSITY_LEVEL level, Args&&... args) const {
    if (level > level_) return;
    static const char* prefix[] = {"ERROR", "WARN", "INFO", "DEBUG"};
    std::ostringstream line;
    line << "[" << prefix[level] << "] duna::" << name_ << "::";
    (line << ... << args) << std::endl;
    const std::string text = line.str();
    sink_ << text;
    for (std::ostream* extra : extra_sinks_) *extra << text;
  }

  void setLogLevel(VERBOSITY_LEVEL level) { level_ = level; }
  void addSink(std::ostream* sink) { extra_sinks_.insert(sink); }

 private:
  std::ostream& sink_;
  VERBOSITY_LEVEL level_;
  std::string name_;
  std::unordered_set<std::ostream*> extra_sinks_;
};

}  // namespace duna
