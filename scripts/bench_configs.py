"""Throughput of the other BASELINE.json configurations (parity-test cases, not the bench.py line):
   configs[1] curve fitting 10 M samples, numerical central differences
   configs[4] camera calibration 50 M observations, numerical Jacobian
   + point2point variants (numerical, fp64 store).  Prints one JSON object per case.
   `--only curve,camera,camera15,user,p2p,fachada` restricts the run to some sections."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moptimizer_0_b200 import capi
from oracle import oracle_py as orc
from tests.common import camera_consts

ONLY = None
if "--only" in sys.argv:
    ONLY = set(sys.argv[sys.argv.index("--only") + 1].split(","))


def section(name):
    return ONLY is None or name in ONLY


ctx = capi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())


def time_pass(store, prob, x, steps=30, warm=100):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    for _ in range(warm):
        ctx.linearize_async(store, prob, x)
    ctx.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        ctx.linearize_async(store, prob, x)
    b.record(stream)
    ctx.synchronize()
    return a.elapsed_time(b) / steps


def report(name, n, bytes_per, ms, extra=None):
    d = {"case": name, "n": n, "ms_per_pass": ms, "Gres_per_s": n / ms / 1e6, "GB_per_s": n * bytes_per / ms / 1e6}
    d.update(extra or {})
    print(json.dumps(d), flush=True)


_lm_warm = [False]


def lm_time(stores, probs, x0, **kw):
    if not _lm_warm[0]:  # first launches of the optimizer kernels (lazy module load) are not part of the rate
        ctx.lm_minimize(stores, probs, x0, max_iterations=1)
        _lm_warm[0] = True
    ctx.synchronize()
    t0 = time.perf_counter()
    r = ctx.lm_minimize(stores, probs, x0, **kw)
    dt = time.perf_counter() - t0
    return r, dt


# ---- curve fitting 10 M --------------------------------------------------------------------------
if section("curve"):
    n = 10_000_000
    st = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
    st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
    for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_FORWARD, "forward"), (capi.JAC_ANALYTICAL, "analytical")):
        for cd, cn in ((capi.F32, "f32"), (capi.F64, "f64")):
            prob = capi.make_problem(capi.MODEL_EXP_CURVE, jac, cd)
            report(f"curve10M_{jn}_{cn}", n, 8, time_pass(st, prob, [0.25, 0.15]))
    prob = capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F64)
    r, dt = lm_time([st], [prob], [0.0, 0.0], max_iterations=50)
    print(json.dumps({"case": "curve10M_lm_central_f64", "status": r.status, "iters": r.executed_iterations,
                      "passes": r.num_passes, "seconds": dt, "lm_iters_per_s": r.executed_iterations / dt,
                      "x": r.x.tolist()}), flush=True)
    st.close()

# ---- camera 50 M ---------------------------------------------------------------------------------
n = 50_000_000
consts = camera_consts()
x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027])
if section("camera"):
    M = consts[:12].reshape(3, 4) @ orc.so3_convert6dof(x_gt) @ consts[12:].reshape(4, 4)
    st = capi.Store(ctx, capi.MODEL_PINHOLE, n, capi.F32)
    st.generate(seed=3, gt=M.reshape(-1), lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5)
    for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_FORWARD, "forward")):
        for cd, cn in ((capi.F32, "f32"), (capi.F64, "f64")):
            prob = capi.make_problem(capi.MODEL_PINHOLE, jac, cd, consts=consts)
            report(f"camera50M_{jn}_{cn}", n, 20, time_pass(st, prob, [0.0] * 6, steps=10, warm=10))
            if cd == capi.F32:  # per-residual float quotient instead of the common-denominator form
                prob = capi.make_problem(capi.MODEL_PINHOLE, jac, cd, consts=consts, flags=capi.FLAG_GENERIC_KERNEL)
                report(f"camera50M_{jn}_{cn}_generic_kernel", n, 20, time_pass(st, prob, [0.0] * 6, steps=10, warm=10))
            else:  # fp64 compute opted into the common-denominator form
                prob = capi.make_problem(capi.MODEL_PINHOLE, jac, cd, consts=consts, flags=capi.FLAG_STABLE_FD)
                report(f"camera50M_{jn}_{cn}_stable_fd", n, 20, time_pass(st, prob, [0.0] * 6, steps=10, warm=10))
    prob = capi.make_problem(capi.MODEL_PINHOLE, capi.JAC_CENTRAL, capi.F64, consts=consts)
    r, dt = lm_time([st], [prob], [0.0] * 6, max_iterations=50)
    print(json.dumps({"case": "camera50M_lm_central_f64", "status": r.status, "iters": r.executed_iterations,
                      "passes": r.num_passes, "seconds": dt, "lm_iters_per_s": r.executed_iterations / dt,
                      "x_err": float(np.max(np.abs(r.x - x_gt)))}), flush=True)
    prob = capi.make_problem(capi.MODEL_PINHOLE, capi.JAC_CENTRAL, capi.F32, consts=consts)
    r, dt = lm_time([st], [prob], [0.0] * 6, max_iterations=50)
    print(json.dumps({"case": "camera50M_lm_central_f32", "status": r.status, "iters": r.executed_iterations,
                      "passes": r.num_passes, "seconds": dt, "lm_iters_per_s": r.executed_iterations / dt,
                      "x_err": float(np.max(np.abs(r.x - x_gt)))}), flush=True)
    st.close()

# ---- the n x n calibration case: pinhole + distortion, P = 15 (BASELINE configs[4]) -----------------------
if section("camera15"):
    x15 = np.concatenate([x_gt, [600.0, 600.0, 320.0, 240.0, 0.05, -0.02, 0.001, -0.001, 0.005]])
    C44 = consts[12:]
    st = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n, capi.F32)
    st.generate(seed=3, gt=x15, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=C44)
    for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_FORWARD, "forward")):
        for cd, cn in ((capi.F32, "f32"), (capi.F64, "f64")):
            prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, jac, cd, consts=C44)
            report(f"camera15_50M_{jn}_{cn}", n, 20, time_pass(st, prob, x15 * 0.999, steps=5, warm=3))
            if cd == capi.F32:
                prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, jac, cd, consts=C44, flags=capi.FLAG_GENERIC_KERNEL)
                report(f"camera15_50M_{jn}_{cn}_generic_kernel", n, 20, time_pass(st, prob, x15 * 0.999, steps=5, warm=3))
            else:
                prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, jac, cd, consts=C44, flags=capi.FLAG_STABLE_FD)
                report(f"camera15_50M_{jn}_{cn}_stable_fd", n, 20, time_pass(st, prob, x15 * 0.999, steps=5, warm=3))
    x0 = x15.copy()
    x0[:6] = 0.0
    x0[6:10] *= 1.02
    x0[10:] = 0.0
    prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F64, consts=C44)
    r, dt = lm_time([st], [prob], x0, max_iterations=50)
    print(json.dumps({"case": "camera15_50M_lm_central_f64", "status": r.status, "iters": r.executed_iterations,
                      "passes": r.num_passes, "seconds": dt, "lm_iters_per_s": r.executed_iterations / dt,
                      "x_err_extrinsics": float(np.max(np.abs(r.x[:6] - x15[:6]))),
                      "x_err_focal_rel": float(np.max(np.abs(r.x[6:8] / x15[6:8] - 1.0)))}), flush=True)
    prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F32, consts=C44)
    r, dt = lm_time([st], [prob], x0, max_iterations=50)
    print(json.dumps({"case": "camera15_50M_lm_central_f32", "status": r.status, "iters": r.executed_iterations,
                      "passes": r.num_passes, "seconds": dt, "lm_iters_per_s": r.executed_iterations / dt,
                      "x_err_extrinsics": float(np.max(np.abs(r.x[:6] - x15[:6]))),
                      "x_err_focal_rel": float(np.max(np.abs(r.x[6:8] / x15[6:8] - 1.0))),
                      "x_err_distortion": float(np.max(np.abs(r.x[10:] - x15[10:])))}), flush=True)
    st.close()

# ---- a user model compiled at run time (NVRTC) next to the builtin it restates --------------------------
if section("user"):
    n = 10_000_000
    SRC = """
    template <typename T> __device__ void mopt_f(const T* s, const T* a, const T* b, T* r) { r[0] = b[0] - exp(fma(s[0], a[0], s[1])); }
    template <typename T> __device__ void mopt_f_df(const T* s, const T* a, const T* b, T* r, T* J) {
      const T ex = exp(fma(s[0], a[0], s[1])); r[0] = b[0] - ex; J[0] = -a[0] * ex; J[1] = -ex; }
    """
    t0 = time.perf_counter()
    um = capi.UserModel(SRC, 2, 1, 1, 1, has_jacobian=True)
    t_compile = time.perf_counter() - t0
    bst = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
    bst.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
    ust = capi.Store(ctx, um.model, n, capi.F32)
    ust.upload(0, bst.download(0, np.float32))
    ust.upload(1, bst.download(1, np.float32))
    for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_ANALYTICAL, "analytical")):
        t0 = time.perf_counter()
        ctx.linearize(ust, capi.make_problem(um.model, jac, capi.F32), [0.25, 0.15])
        t_first = time.perf_counter() - t0
        mb = time_pass(bst, capi.make_problem(capi.MODEL_EXP_CURVE, jac, capi.F32), [0.25, 0.15])
        mu = time_pass(ust, capi.make_problem(um.model, jac, capi.F32), [0.25, 0.15])
        report(f"user_curve10M_{jn}_f32", n, 8, mu, {"builtin_ms_per_pass": mb, "first_call_s_incl_nvrtc": t_first,
                                                     "initial_compile_s": t_compile})
    bst.close(); ust.close()

# ---- point2point variants ------------------------------------------------------------------------
if section("p2p"):
    n = 100_000_000
    X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
    st.generate(seed=2, gt=X_GT, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    for jac, jn in ((capi.JAC_ANALYTICAL, "analytical"), (capi.JAC_FORWARD, "forward"), (capi.JAC_CENTRAL, "central")):
        for cd, cn in ((capi.F32, "f32"), (capi.F64, "f64")):
            prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, cd, loss=capi.LOSS_HUBER, loss_param=0.05)
            report(f"p2p100M_f32store_{jn}_{cn}", n, 24, time_pass(st, prob, [0.0] * 6, steps=10, warm=10))
            if jac != capi.JAC_ANALYTICAL:  # per-residual difference quotient instead of the moment kernel
                prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, cd, loss=capi.LOSS_HUBER, loss_param=0.05,
                                         flags=capi.FLAG_GENERIC_KERNEL)
                report(f"p2p100M_f32store_{jn}_{cn}_generic_kernel", n, 24, time_pass(st, prob, [0.0] * 6, steps=10, warm=10))
    st.close()
    n = 50_000_000
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F64)
    st.generate(seed=2, gt=X_GT, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F64, loss=capi.LOSS_HUBER, loss_param=0.05)
    report("p2p50M_f64store_analytical_f64", n, 48, time_pass(st, prob, [0.0] * 6, steps=10, warm=10))
    st.close()

# ---- small problem LM rate: fachada (29 310 points), fp64 -----------------------------------------
if section("fachada"):
    from tests.common import fachada
    src, tgt, _, _ = fachada()
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F64)
    st.upload(0, src); st.upload(1, tgt)
    for jac, jn in ((capi.JAC_ANALYTICAL, "analytical"), (capi.JAC_FORWARD, "forward")):
        prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64)
        ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50)
        best = 1e9
        for _ in range(20):
            r, dt = lm_time([st], [prob], [0.0] * 6, max_iterations=50)
            best = min(best, dt)
        print(json.dumps({"case": f"fachada_lm_{jn}_f64", "status": r.status, "iters": r.executed_iterations,
                          "passes": r.num_passes, "best_seconds": best, "lm_iters_per_s": r.executed_iterations / best}), flush=True)
        report(f"fachada_pass_{jn}_f64", src.shape[0], 48, time_pass(st, prob, [0.0] * 6, steps=200, warm=50))
    st.close()
ctx.close()
