// Mirrors tst/test_main.cpp: one binary, every TEST registered by the other translation units.
#include "mini_test.h"

#include "moptimizer/device/context.h"

moptimizer::device::Context::Ptr g_ctx;

int main(int argc, char** argv) {
  const char* filter = argc > 1 ? argv[1] : nullptr;
  try {
    g_ctx = moptimizer::device::Context::create(0);
  } catch (const std::exception& e) {
    std::printf("no usable CUDA device: %s\n", e.what());
    return 77;
  }
  const int failed = mini::run_all(filter);
  g_ctx.reset();
  return failed ? 1 : 0;
}
