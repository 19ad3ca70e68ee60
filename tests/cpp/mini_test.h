// Tiny gtest-shaped harness (GoogleTest is not in this image) so the C++ tests read like the reference's.
#pragma once

#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace mini {
struct Case {
  std::string name;
  std::function<void()> fn;
};
inline std::vector<Case>& registry() {
  static std::vector<Case> r;
  return r;
}
inline int& failures() {
  static int f = 0;
  return f;
}
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<void()> fn) {
    registry().push_back({std::string(suite) + "." + name, std::move(fn)});
  }
};
inline int run_all(const char* filter) {
  int ran = 0, failed_cases = 0;
  for (auto& c : registry()) {
    if (filter && c.name.find(filter) == std::string::npos) continue;
    const int before = failures();
    std::printf("[ RUN      ] %s\n", c.name.c_str());
    try {
      c.fn();
    } catch (const std::exception& e) {
      std::printf("  unexpected exception: %s\n", e.what());
      ++failures();
    }
    const bool ok = failures() == before;
    std::printf("[ %s ] %s\n", ok ? "      OK" : " FAILED ", c.name.c_str());
    if (!ok) ++failed_cases;
    ++ran;
  }
  std::printf("%d tests ran, %d failed\n", ran, failed_cases);
  return failed_cases;
}
}  // namespace mini

#define TEST(suite, name)                                                        \
  static void suite##_##name##_body();                                           \
  static mini::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
  static void suite##_##name##_body()

#define EXPECT_NEAR(a, b, tol)                                                                          \
  do {                                                                                                  \
    const double _a = double(a), _b = double(b), _t = double(tol);                                      \
    if (!(std::fabs(_a - _b) <= _t)) {                                                                  \
      std::printf("  %s:%d EXPECT_NEAR(%s, %s, %s): %.12g vs %.12g\n", __FILE__, __LINE__, #a, #b, #tol, _a, _b); \
      ++mini::failures();                                                                               \
    }                                                                                                   \
  } while (0)
#define EXPECT_TRUE(c)                                                             \
  do {                                                                             \
    if (!(c)) {                                                                    \
      std::printf("  %s:%d EXPECT_TRUE(%s)\n", __FILE__, __LINE__, #c);            \
      ++mini::failures();                                                          \
    }                                                                              \
  } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_THROW(stmt, ex)                                                     \
  do {                                                                             \
    bool _caught = false;                                                          \
    try {                                                                          \
      stmt;                                                                        \
    } catch (const ex&) {                                                          \
      _caught = true;                                                              \
    } catch (...) {                                                                \
    }                                                                              \
    if (!_caught) {                                                                \
      std::printf("  %s:%d EXPECT_THROW(%s, %s)\n", __FILE__, __LINE__, #stmt, #ex); \
      ++mini::failures();                                                          \
    }                                                                              \
  } while (0)
