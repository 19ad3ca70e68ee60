// Mirrors include/moptimizer/levenberg_marquadt_dyn.h:8-50 + src/levenberg_marquadt_dyn.cpp.  minimize() hands
// the whole loop (passes, damped LDL^T solve, gain ratio, accept/reject) to the device through
// mopt_lm_minimize; the iteration trace comes back once, at the end, and is logged at DEBUG like the
// reference's per-iteration lines (levenberg_marquadt_dyn.cpp:41,72-75,94-95).
#pragma once

#include <vector>

#include "mopt_capi.h"
#include "moptimizer/delta.h"
#include "moptimizer/optimizer.h"

namespace moptimizer {

template <class Scalar>
class LevenbergMarquadtDynamic : public Optimizer<Scalar> {
 public:
  explicit LevenbergMarquadtDynamic(int num_parameters)
      : num_parameters_(num_parameters), lm_init_lambda_factor_(Scalar(1e-9)), lm_lambda_(Scalar(-1)),
        lm_max_iterations_(3) {}
  ~LevenbergMarquadtDynamic() override = default;

  /// levenberg_marquadt_dyn.cpp:29-31 — a stub in the reference as well.
  OptimizationStatus step(Scalar*) override { return OptimizationStatus::NUMERIC_ERROR; }

  OptimizationStatus minimize(Scalar* x0) override {
    this->checkCosts();
    prepare(x0);
    std::vector<mopt_problem> problems;
    std::vector<device::Store::Ptr> stores;
    std::vector<mopt_store*> raw;
    this->gatherDeviceProblems(x0, &problems, &stores, &raw);  // cost->update(x0) (:54) + device descriptions
    const int n = int(raw.size());
    mopt_lm_options opt;
    mopt_lm_default_options(&opt);
    opt.max_iterations = int(maximum_iterations_);
    opt.lm_max_iterations = int(lm_max_iterations_);
    opt.lambda_factor = double(lm_init_lambda_factor_);
    opt.scalar_dtype = device::dtypeOf<Scalar>();
    opt.speculative = speculative_ ? 1 : 0;
    opt.flags = stagnation_stop_ ? MOPT_LM_STAGNATION_STOP : 0;
    double xd[MOPT_MAX_PARAMETERS] = {0};
    for (int i = 0; i < num_parameters_; ++i) xd[i] = double(x0[i]);
    report_.reset(new mopt_lm_report());
    device::check(mopt_lm_minimize(stores[0]->context()->get(), n, raw.data(), problems.data(), &opt, xd, report_.get()),
                  "mopt_lm_minimize");
    for (int i = 0; i < num_parameters_; ++i) x0[i] = Scalar(xd[i]);
    executed_iterations_ = unsigned(report_->executed_iterations);
    for (int t = 0; t < report_->num_trials; ++t) {
      const mopt_lm_trial& tr = report_->trials[t];
      logger_->log(duna::Logger::L_DEBUG, "Internal Iteration --- : ", tr.outer_iteration, ':', tr.k + 1, '/',
                   lm_max_iterations_, ' ', tr.y0, ' ', tr.yi, ' ', tr.rho, ' ', tr.lambda, ' ', tr.nu,
                   tr.accepted ? " accepted" : " rejected");
    }
    if (report_->status == MOPT_NUMERIC_ERROR) logger_->log(duna::Logger::L_ERROR, "Numeric Error!");
    return static_cast<OptimizationStatus>(report_->status);
  }

  inline void setLogger(std::shared_ptr<duna::Logger> logger) { logger_ = logger; }
  inline unsigned int getLevenbergMarquadtIterations() const { return lm_max_iterations_; }
  inline void setLevenbergMarquadtIterations(int max_iterations) { lm_max_iterations_ = max_iterations; }

  /// New: 1 (default) fuses cost and linearization of every trial point into one pass; 0 keeps the reference's
  /// pass order (linearize at x, then cost at each trial).  Same results either way.
  inline void setSpeculativeLinearization(bool on) { speculative_ = on; }
  /// New: on (default) ends with SMALL_DELTA once an accepted step is below isDeltaSmall's threshold and leaves the
  /// cost bit-for-bit unchanged; off is the literal loop (rho = 0 steps until the iterations run out, see
  /// MOPT_LM_STAGNATION_STOP in mopt_capi.h).
  inline void setStagnationStop(bool on) { stagnation_stop_ = on; }
  /// New: the device iteration trace of the last minimize() (nullptr before the first call).
  const mopt_lm_report* report() const { return report_.get(); }

 protected:
  int num_parameters_;
  bool hasConverged() override { return false; }
  void prepare(Scalar*) override {  // levenberg_marquadt_dyn.cpp:15-26
    lm_init_lambda_factor_ = Scalar(1e-9);
    lm_lambda_ = Scalar(-1.0);
  }

  using Optimizer<Scalar>::costs_;
  using Optimizer<Scalar>::maximum_iterations_;
  using Optimizer<Scalar>::executed_iterations_;
  using Optimizer<Scalar>::logger_;

  Scalar lm_init_lambda_factor_;
  Scalar lm_lambda_;
  unsigned int lm_max_iterations_;
  bool speculative_ = true;
  bool stagnation_stop_ = true;
  std::unique_ptr<mopt_lm_report> report_;
};

}  // namespace moptimizer
