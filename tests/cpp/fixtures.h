// Loads tests/golden/reference_fixtures.txt (+ fachada_xyz.f64): the data the reference's tests use.
#pragma once

#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef TEST_DATA_PATH
#error "TEST_DATA_PATH NOT DEFINED!"
#endif

inline const std::map<std::string, std::vector<double>>& fixtures() {
  static std::map<std::string, std::vector<double>> fx;
  if (fx.empty()) {
    std::ifstream in(std::string(TEST_DATA_PATH) + "/reference_fixtures.txt");
    if (!in) throw std::runtime_error("cannot open reference_fixtures.txt");
    std::string line;
    while (std::getline(in, line)) {
      std::istringstream ss(line);
      std::string name;
      size_t n;
      ss >> name >> n;
      std::vector<double> v(n);
      for (auto& x : v) ss >> x;
      fx[name] = v;
    }
  }
  return fx;
}
inline const std::vector<double>& fx(const char* k) { return fixtures().at(k); }

inline std::vector<double> load_fachada() {
  const std::string path = std::string(TEST_DATA_PATH) + "/fachada_xyz.f64";
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("not a file! exiting");
  std::fseek(f, 0, SEEK_END);
  const long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<double> v(bytes / sizeof(double));
  if (std::fread(v.data(), sizeof(double), v.size(), f) != v.size()) throw std::runtime_error("short read");
  std::fclose(f);
  return v;
}
