// The snippet of INTEGRATION.md §1 as a program: rigid registration of two clouds through the C++ mirror of the
// reference API (same class names and call sequence as tst/point2point.cpp:142-160), whole LM loop on the GPU.
//
//   g++ -std=c++17 -Iinclude examples/point2point_registration.cpp -Lmoptimizer_0_b200 -lmopt_b200
//       -Wl,-rpath,$PWD/moptimizer_0_b200 -o /tmp/p2p_example && /tmp/p2p_example [n]
//
// Exits 77 with a message when no CUDA device is usable: there is no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <moptimizer/cost_function_analytical_dyn.h>
#include <moptimizer/levenberg_marquadt_dyn.h>
#include <moptimizer/loss_function/huber.h>
#include <moptimizer/device/models.h>
#include <moptimizer/so3.h>

int main(int argc, char** argv) {
  const long n = argc > 1 ? std::atol(argv[1]) : 1000000;
  // synthetic clouds: tgt = T_gt src + noise, 5 % outliers (what bench.py generates on the device)
  const double x_gt[6] = {0.5, -0.3, 0.2, 0.10, -0.05, 0.08};
  double T[16];
  so3::convert6DOFParameterToMatrix<double>(x_gt, T);
  std::mt19937_64 rng(2);
  std::uniform_real_distribution<double> box(0.0, 10.0), out(-1.0, 1.0), u01(0.0, 1.0);
  std::normal_distribution<double> noise(0.0, 0.01);
  std::vector<double> src(3 * n), tgt(3 * n);
  for (long i = 0; i < n; ++i) {
    double* p = &src[3 * i];
    for (int k = 0; k < 3; ++k) p[k] = box(rng);
    const bool outlier = u01(rng) < 0.05;
    for (int k = 0; k < 3; ++k)
      tgt[3 * i + k] = T[k * 4] * p[0] + T[k * 4 + 1] * p[1] + T[k * 4 + 2] * p[2] + T[k * 4 + 3] + noise(rng) +
                       (outlier ? out(rng) : 0.0);
  }
  try {
    auto ctx = moptimizer::device::Context::create(0);
    // fp32 planar streams in HBM (24 B per correspondence), fp64 LM arithmetic: the benchmark configuration
    auto model = std::make_shared<moptimizer::device::Point2Point<double>>(ctx, src.data(), tgt.data(), n,
                                                                          MOPT_P2P_EXACT, MOPT_F32);
    moptimizer::CostFunctionAnalyticalDynamic<double> cost(model, 6, 3, int(n));
    cost.setLossFunction(std::make_shared<moptimizer::loss::Huber<double>>(0.05));
    moptimizer::LevenbergMarquadtDynamic<double> lm(6);
    lm.setMaximumIterations(50);
    lm.addCost(&cost);
    double x0[6] = {0, 0, 0, 0, 0, 0};
    const auto status = lm.minimize(x0);
    double err = 0;
    for (int i = 0; i < 6; ++i) err = std::fmax(err, std::fabs(x0[i] - x_gt[i]));
    std::printf("status %d after %d iterations, x = (%.5f %.5f %.5f %.5f %.5f %.5f), max |x - x_gt| = %.2e\n", int(status),
                lm.getExecutedIterations(), x0[0], x0[1], x0[2], x0[3], x0[4], x0[5], err);
    return err < 1e-3 ? 0 : 1;
  } catch (const moptimizer::Exception& e) {
    std::printf("moptimizer::Exception: %s\n", e.what());
    return 77;
  }
}
