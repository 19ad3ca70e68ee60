// ORACLE — TEST INFRASTRUCTURE ONLY (see moptimizer_oracle.hpp header).
// C interface over the restated reference so that pytest (ctypes) and bench.py's
// cpu_baseline / `--impl reference` legs can drive it.  Never loaded by the product.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "moptimizer_oracle.hpp"

using namespace oracle;

extern "C" {

enum { ORC_P2P = 0, ORC_EXP_CURVE = 1, ORC_MICHAELIS_MENTEN = 2, ORC_PINHOLE = 3, ORC_POWELL = 4,
       ORC_POINT_DIST = 5, ORC_PINHOLE_DISTORT = 6, ORC_P2P_ICP = 7 };
enum { ORC_LOSS_NONE = 0, ORC_LOSS_GM = 1, ORC_LOSS_HUBER = 2 };

struct orc_cost {
  int model;
  int variant;  // P2PJacobian for ORC_P2P
  int P, O, n;
  int jac_mode;  // JacobianMode
  int loss;
  double loss_param;
  const double* cov;     // O*O column-major, or NULL for identity
  const void* a;         // data stream A (model dependent), dtype `data_f32 ? float : double`
  const void* b;         // data stream B
  int data_f32;          // 1 if a/b are float arrays
  const double* consts;  // pinhole: K (12, row-major) then C (16, row-major)
  int cost_threads;      // threads for computeCost's parallel reduce (>=1)
  int float_carry;       // emulate the oneTBB float-identity quirk in computeCost
  int manifold;          // Manifold: 0 additive (reference), 1 left SO(3) perturbation of x[3..5]
  const void* target;    // ORC_P2P_ICP: fixed target cloud (xyz AoS, same dtype as a/b)
  int target_m;
  double max_dist;       // ORC_P2P_ICP: maximum correspondence distance
  const double* update_x;  // ORC_P2P_ICP: parameters at which cost->update(x) ran before linearize / computeCost
                           // (NULL: at the evaluation point)
  int lin_threads;         // threads for Cost::linearize inside orc_lm_minimize (<= 1: the reference's serial loop)
};

}  // extern "C"

namespace {

template <class S>
struct Holder {
  std::vector<S> a_store, b_store, t_store;
  const S* a = nullptr;
  const S* b = nullptr;
  const S* target = nullptr;
  Cost<S> cost;
};

template <class S>
const S* adopt(const void* p, bool is_f32, size_t count, std::vector<S>& store) {
  if (!p || count == 0) return nullptr;
  if (is_f32 == std::is_same<S, float>::value) return static_cast<const S*>(p);
  store.resize(count);
  if (is_f32) {
    const float* f = static_cast<const float*>(p);
    for (size_t i = 0; i < count; ++i) store[i] = S(f[i]);
  } else {
    const double* d = static_cast<const double*>(p);
    for (size_t i = 0; i < count; ++i) store[i] = S(d[i]);
  }
  return store.data();
}

template <class S>
std::unique_ptr<Holder<S>> build(const orc_cost& c) {
  auto h = std::make_unique<Holder<S>>();
  size_t na = 0, nb = 0;
  switch (c.model) {
    case ORC_P2P:
    case ORC_POINT_DIST: na = nb = size_t(c.n) * 3; break;
    case ORC_P2P_ICP: na = size_t(c.n) * 3; nb = 0; break;
    case ORC_EXP_CURVE:
    case ORC_MICHAELIS_MENTEN: na = nb = size_t(c.n); break;
    case ORC_PINHOLE:
    case ORC_PINHOLE_DISTORT: na = size_t(c.n) * 3; nb = size_t(c.n) * 2; break;
    default: break;
  }
  h->a = adopt<S>(c.a, c.data_f32 != 0, na, h->a_store);
  h->b = adopt<S>(c.b, c.data_f32 != 0, nb, h->b_store);
  Cost<S>& k = h->cost;
  switch (c.model) {
    case ORC_P2P: k.model = std::make_shared<Point2Point<S>>(h->a, h->b, c.variant); break;
    case ORC_P2P_ICP:
      h->target = adopt<S>(c.target, c.data_f32 != 0, size_t(c.target_m) * 3, h->t_store);
      k.model = std::make_shared<Point2PointICP<S>>(h->a, c.n, h->target, c.target_m, S(c.max_dist), c.variant);
      break;
    case ORC_POINT_DIST: k.model = std::make_shared<PointDist<S>>(h->a, h->b); break;
    case ORC_EXP_CURVE: k.model = std::make_shared<ExpCurve<S>>(h->a, h->b); break;
    case ORC_MICHAELIS_MENTEN: k.model = std::make_shared<MichaelisMenten<S>>(h->a, h->b); break;
    case ORC_POWELL: k.model = std::make_shared<Powell<S>>(); break;
    case ORC_PINHOLE_DISTORT: {  // consts = C (16, row-major)
      S C[16];
      for (int i = 0; i < 16; ++i) C[i] = S(c.consts[i]);
      k.model = std::make_shared<PinholeDistort<S>>(h->a, h->b, C);
      break;
    }
    case ORC_PINHOLE: {
      S K[12], C[16];
      for (int i = 0; i < 12; ++i) K[i] = S(c.consts[i]);
      for (int i = 0; i < 16; ++i) C[i] = S(c.consts[12 + i]);
      k.model = std::make_shared<Pinhole<S>>(h->a, h->b, K, C);
      break;
    }
    default: return nullptr;
  }
  switch (c.loss) {
    case ORC_LOSS_NONE: k.loss = std::make_shared<NoLoss<S>>(); break;
    case ORC_LOSS_GM: k.loss = std::make_shared<GemanMcClure<S>>(S(c.loss_param)); break;
    case ORC_LOSS_HUBER: k.loss = std::make_shared<Huber<S>>(S(c.loss_param)); break;
    default: return nullptr;
  }
  k.P = c.P;
  k.O = c.O;
  k.n = c.n;
  k.jac_mode = c.jac_mode;
  k.cost_threads = c.cost_threads < 1 ? 1 : c.cost_threads;
  k.float_carry = c.float_carry != 0;
  k.manifold = c.manifold;
  k.lin_threads = c.lin_threads < 1 ? 1 : c.lin_threads;
  k.C.assign(size_t(c.O) * c.O, S(0));
  for (int i = 0; i < c.O; ++i) k.C[i + size_t(i) * c.O] = S(1);
  if (c.cov)
    for (int i = 0; i < c.O * c.O; ++i) k.C[i] = S(c.cov[i]);
  if (c.P > CostComputation<S>::kMaxP || c.O > CostComputation<S>::kMaxO) return nullptr;
  return h;
}

// cost->update(x) (cost_function.h:44) for models that implement it, before a stand-alone evaluation
template <class S>
void run_update(const orc_cost* c, Cost<S>& cost, const std::vector<S>& x_eval) {
  if (c->model != ORC_P2P_ICP) return;
  std::vector<S> xu(x_eval);
  if (c->update_x)
    for (int i = 0; i < c->P; ++i) xu[i] = S(c->update_x[i]);
  cost.model->update(xu.data());
}

template <class S>
int linearize_t(const orc_cost* c, const double* x, double* H, double* b, double* sum,
                int nthreads) {
  auto h = build<S>(*c);
  if (!h) return 1;
  const int P = c->P;
  std::vector<S> xs(P), Hs(size_t(P) * P), bs(P);
  for (int i = 0; i < P; ++i) xs[i] = S(x[i]);
  S s;
  run_update<S>(c, h->cost, xs);
  if (nthreads > 1) {
    CostComputation<S> cc(P, c->O);
    cc.manifold_ = c->manifold;
    s = cc.parallelLinearize(xs.data(), h->cost.C.data(), *h->cost.loss, Hs.data(), bs.data(),
                             *h->cost.model, c->n, c->jac_mode, nthreads);
  } else {
    s = h->cost.linearize(xs.data(), Hs.data(), bs.data());
  }
  for (int i = 0; i < P * P; ++i) H[i] = double(Hs[i]);
  for (int i = 0; i < P; ++i) b[i] = double(bs[i]);
  *sum = double(s);
  return 0;
}

template <class S>
int cost_t(const orc_cost* c, const double* x, double* sum, int parallel) {
  auto h = build<S>(*c);
  if (!h) return 1;
  std::vector<S> xs(c->P > 0 ? c->P : 1);
  for (int i = 0; i < c->P; ++i) xs[i] = S(x[i]);
  run_update<S>(c, h->cost, xs);
  CostComputation<S> cc(c->P, c->O);
  S s = parallel ? cc.parallelComputeCost(xs.data(), *h->cost.model, c->n, h->cost.cost_threads,
                                          h->cost.float_carry)
                 : cc.computeCost(xs.data(), *h->cost.model, c->n);
  *sum = double(s);
  return 0;
}

template <class S>
int lm_t(const orc_cost* cs, int ncosts, int P, int max_it, int lm_it, double* x, int* status,
         int* executed, double* trace, int max_trace, int* ntrace, int stagnation_stop) {
  std::vector<std::unique_ptr<Holder<S>>> hs;
  std::vector<Cost<S>*> costs;
  for (int i = 0; i < ncosts; ++i) {
    hs.push_back(build<S>(cs[i]));
    if (!hs.back()) return 1;
    costs.push_back(&hs.back()->cost);
  }
  std::vector<S> xs(P);
  for (int i = 0; i < P; ++i) xs[i] = S(x[i]);
  std::vector<TraceEntry> tr;
  Status st = lm_minimize<S>(costs, P, max_it, lm_it, xs.data(), executed, &tr, stagnation_stop != 0);
  for (int i = 0; i < P; ++i) x[i] = double(xs[i]);
  *status = int(st);
  int nt = 0;
  for (auto& e : tr) {
    if (nt >= max_trace) break;
    double* t = trace + size_t(nt) * 8;
    t[0] = e.outer_it; t[1] = e.k; t[2] = e.y0; t[3] = e.yi; t[4] = e.rho; t[5] = e.lambda;
    t[6] = e.nu; t[7] = e.accepted;
    ++nt;
  }
  *ntrace = nt;
  return 0;
}

}  // namespace

extern "C" {

// scalar: 0 = float, 1 = double (the reference's two instantiations, src/*.cpp tails).
int orc_linearize(const orc_cost* c, int scalar, const double* x, double* H, double* b,
                  double* sum, int nthreads) {
  return scalar ? linearize_t<double>(c, x, H, b, sum, nthreads)
                : linearize_t<float>(c, x, H, b, sum, nthreads);
}

int orc_compute_cost(const orc_cost* c, int scalar, const double* x, double* sum, int parallel) {
  return scalar ? cost_t<double>(c, x, sum, parallel) : cost_t<float>(c, x, sum, parallel);
}

// trace: max_trace rows of 8 doubles {outer_it, k, y0, yi, rho, lambda, nu, accepted}.
// stagnation_stop: 0 = the reference's loop, 1 = with the product's MOPT_LM_STAGNATION_STOP rule restated.
int orc_lm_minimize(const orc_cost* costs, int ncosts, int scalar, int P, int max_it, int lm_it,
                    double* x, int* status, int* executed, double* trace, int max_trace,
                    int* ntrace, int stagnation_stop) {
  return scalar ? lm_t<double>(costs, ncosts, P, max_it, lm_it, x, status, executed, trace,
                               max_trace, ntrace, stagnation_stop)
                : lm_t<float>(costs, ncosts, P, max_it, lm_it, x, status, executed, trace,
                              max_trace, ntrace, stagnation_stop);
}

int orc_ldlt_solve(int n, const double* A, const double* rhs, double* out) {
  ldlt_solve<double>(n, A, rhs, out);
  return 0;
}

int orc_so3_convert6dof(const double* x, double* T16_rowmajor) {
  so3_convert6dof<double>(x, T16_rowmajor);
  return 0;
}

int orc_so3_left_jacobian_full(const double* w, double* J9_rowmajor) {
  so3_left_jacobian_full<double>(w, J9_rowmajor);
  return 0;
}

// Timed loops for bench.py (cpu_baseline / --impl reference): runs `reps` linearizations
// and returns the elapsed seconds of the loop only (model construction excluded).
double orc_time_linearize(const orc_cost* c, int scalar, const double* x, int nthreads, int reps,
                          double* H, double* b, double* sum) {
  if (!scalar) return -1.0;
  auto h = build<double>(*c);
  if (!h) return -1.0;
  const int P = c->P;
  std::vector<double> Hs(size_t(P) * P), bs(P);
  double s = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; ++r) {
    if (nthreads > 1) {
      CostComputation<double> cc(P, c->O);
      cc.manifold_ = c->manifold;
      s = cc.parallelLinearize(x, h->cost.C.data(), *h->cost.loss, Hs.data(), bs.data(),
                               *h->cost.model, c->n, c->jac_mode, nthreads);
    } else {
      s = h->cost.linearize(x, Hs.data(), bs.data());
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  std::memcpy(H, Hs.data(), sizeof(double) * Hs.size());
  std::memcpy(b, bs.data(), sizeof(double) * bs.size());
  *sum = s;
  return std::chrono::duration<double>(t1 - t0).count();
}

int orc_hardware_concurrency() { return int(std::thread::hardware_concurrency()); }

// ---- the bench workload on the host ---------------------------------------------------------------------------
// Host restatement of the product's counter-based synthetic generator for the point2point model
// (moptimizer_0_b200/csrc/mopt_store.cu generate_p2p_kernel; same hash, same explicitly rounded operations in the
// same order, fmaf where the device uses __fmaf_rn): element i of the output equals element i of a device store
// generated with the same description, bit for bit (asserted on a GPU in tests/test_gpu_parity.py).  This is how
// `bench.py --impl reference` times the CPU path on the SAME 100 M correspondences without touching a GPU.
namespace {
inline uint64_t gen_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline float gen_u01(uint64_t seed, int64_t index, int lane) {
  const uint64_t h = gen_splitmix64(seed ^ gen_splitmix64(uint64_t(index) * 16ull + uint64_t(lane)));
  return float(uint32_t(h >> 40)) * (1.0f / 16777216.0f);
}
inline float gen_normal(uint64_t seed, int64_t index, int lane0) {
  const float s = (gen_u01(seed, index, lane0) + gen_u01(seed, index, lane0 + 1)) +
                  (gen_u01(seed, index, lane0 + 2) + gen_u01(seed, index, lane0 + 3));
  return (s + -2.0f) * 1.7320508f;
}
}  // namespace

// x_gt = [t, omega] (6), lo / hi = source box corners (3 each).  src_xyz / tgt_xyz: n x 3 AoS doubles (each value
// is exactly the float the device store holds).  nthreads >= 1.
int orc_generate_p2p(uint64_t seed, int64_t first_index, int64_t n, const double* x_gt, const double* lo,
                     const double* hi, double sigma_d, double outlier_fraction_d, double outlier_range_d,
                     double* src_xyz, double* tgt_xyz, int nthreads) {
  if (!x_gt || !lo || !hi || !src_xyz || !tgt_xyz || n < 0) return 1;
  // rodrigues_host of mopt_store.cu (guard n > 0), then rounded to float like SynthDev::gt
  float gt[12], flo[3], fhi[3];
  {
    const double* w = x_gt + 3;
    const double nn = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (nn > 0.0) {
      const double a[3] = {w[0] / nn, w[1] / nn, w[2] / nn};
      const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
      const double sn = std::sin(nn), cs = std::cos(nn);
      for (int r = 0; r < 3; ++r)
        for (int col = 0; col < 3; ++col) {
          double kk = 0;
          for (int k = 0; k < 3; ++k) kk += K[r * 3 + k] * K[k * 3 + col];
          R[r * 3 + col] += sn * K[r * 3 + col] + (1.0 - cs) * kk;
        }
    }
    for (int i = 0; i < 9; ++i) gt[i] = float(R[i]);
    for (int i = 0; i < 3; ++i) gt[9 + i] = float(x_gt[i]);
    for (int k = 0; k < 3; ++k) { flo[k] = float(lo[k]); fhi[k] = float(hi[k]); }
  }
  const float sigma = float(sigma_d), ofrac = float(outlier_fraction_d), orange = float(outlier_range_d);
  if (nthreads < 1) nthreads = 1;
  auto body = [&](int t) {
    const int64_t i0 = n * t / nthreads, i1 = n * (t + 1) / nthreads;
    for (int64_t i = i0; i < i1; ++i) {
      const int64_t gi = first_index + i;
      float p[3];
      for (int k = 0; k < 3; ++k) p[k] = std::fmaf(fhi[k] + -flo[k], gen_u01(seed, gi, k), flo[k]);
      const bool outlier = gen_u01(seed, gi, 15) < ofrac;
      for (int k = 0; k < 3; ++k) {
        float v = std::fmaf(gt[k * 3 + 2], p[2], std::fmaf(gt[k * 3 + 1], p[1], std::fmaf(gt[k * 3 + 0], p[0], gt[9 + k])));
        if (sigma > 0.0f) v = std::fmaf(sigma, gen_normal(seed, gi, 3 + 4 * k), v);
        if (outlier) v = std::fmaf(orange, std::fmaf(2.0f, gen_u01(seed, gi ^ 0x5bd1e995, k), -1.0f), v);
        tgt_xyz[i * 3 + k] = double(v);
        src_xyz[i * 3 + k] = double(p[k]);
      }
    }
  };
  if (nthreads == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(body, t);
    for (auto& x : th) x.join();
  }
  return 0;
}

}  // extern "C"
