"""User-defined device models (include/mopt_capi.h mopt_user_model_compile): CUDA C++ source of a
model's f / f_df / setup compiled at run time (NVRTC, sm_100a) into the library's own pass kernels.

CPU tests: the source compiles without a GPU, compile errors come back with the compiler log, bad shapes are
rejected.  GPU tests: a user model that restates a builtin model must give BIT-IDENTICAL results to the builtin
(same kernel template, same arithmetic), match the oracle to the suite's tolerances (H, b <= 1e-10 with fp64
compute, <= 1e-5 with fp32), and reproduce the LM trace; models the library does not ship (Gaussian peak,
degree-9 polynomial with P = 10 -> the wide kernel) are checked against a numpy restatement of
computeHessian / computeHessianNumerical (include/moptimizer/linearization.h:65-158)."""
import numpy as np
import pytest

from tests.common import FX, fachada, rel_err

EXP_CURVE_SRC = r"""
// tst/curve_fitting.cpp:86-93 restated as a user model: r = y - exp(x0 t + x1)
template <typename T> __device__ void mopt_f(const T* s, const T* a, const T* b, T* r) {
  r[0] = b[0] - exp(fma(s[0], a[0], s[1]));
}
template <typename T> __device__ void mopt_f_df(const T* s, const T* a, const T* b, T* r, T* J) {
  const T ex = exp(fma(s[0], a[0], s[1]));
  r[0] = b[0] - ex;
  J[0] = -a[0] * ex;
  J[1] = -ex;
}
"""

P2P_SRC = r"""
// tst/point2point.cpp:32-44 restated: setup builds T(x) with so3::Exp, f = R p + t - q
__device__ void mopt_setup(const double* x, const double* consts, double* s) {
  const double w[3] = {x[3], x[4], x[5]};
  if (consts[0] != 0.0) mopt::so3_exp_dev<float>(w, s); else mopt::so3_exp_dev<double>(w, s);
  s[9] = x[0]; s[10] = x[1]; s[11] = x[2];
}
template <typename T> __device__ void mopt_f(const T* s, const T* a, const T* b, T* r) {
  for (int k = 0; k < 3; ++k)
    r[k] = (fma(s[k * 3 + 0], a[0], fma(s[k * 3 + 1], a[1], s[k * 3 + 2] * a[2])) + s[9 + k]) - b[k];
}
"""

GAUSS_SRC = r"""
// r = y - A exp(-(t - mu)^2 / (2 sigma^2)),  x = (A, mu, sigma)
template <typename T> __device__ void mopt_f(const T* x, const T* a, const T* b, T* r) {
  const T d = a[0] - x[1];
  r[0] = b[0] - x[0] * exp(-(d * d) / (T(2) * x[2] * x[2]));
}
template <typename T> __device__ void mopt_f_df(const T* x, const T* a, const T* b, T* r, T* J) {
  const T d = a[0] - x[1];
  const T g = exp(-(d * d) / (T(2) * x[2] * x[2]));
  r[0] = b[0] - x[0] * g;
  J[0] = -g;
  J[1] = -x[0] * g * d / (x[2] * x[2]);
  J[2] = -x[0] * g * d * d / (x[2] * x[2] * x[2]);
}
"""

POLY_SRC = r"""
// two outputs sharing ten coefficients: r0 = y0 - sum_k x_k t^k, r1 = y1 - u * sum_k x_k t^k   (a = (t, u))
template <typename T> __device__ void mopt_f(const T* x, const T* a, const T* b, T* r) {
  T acc = x[9];
  for (int k = 8; k >= 0; --k) acc = fma(acc, a[0], x[k]);
  r[0] = b[0] - acc;
  r[1] = b[1] - a[1] * acc;
}
template <typename T> __device__ void mopt_f_df(const T* x, const T* a, const T* b, T* r, T* J) {
  mopt_f<T>(x, a, b, r);
  T p = T(1);
  for (int k = 0; k < 10; ++k) {
    J[k] = -p;
    J[10 + k] = -a[1] * p;
    p *= a[0];
  }
}
"""


# ------------------------------------------------------------------------------ CPU ----
def test_user_model_compiles_without_gpu():
    from moptimizer_0_b200 import capi
    m = capi.UserModel(EXP_CURVE_SRC, 2, 1, 1, 1, has_jacobian=True)
    assert m.model >= capi.MODEL_USER_BASE
    w = capi.UserModel(POLY_SRC, 10, 2, 2, 2, has_jacobian=True)  # packed size 66 > 32: wide kernel
    s = capi.UserModel(P2P_SRC, 6, 3, 3, 3, set_size=12, rot_offset=3)
    assert len({m.model, w.model, s.model}) == 3
    for u in (m, w, s):
        u.release()


def test_user_model_compile_error_carries_log():
    from moptimizer_0_b200 import capi
    bad = "template <typename T> __device__ void mopt_f(const T* s, const T* a, const T* b, T* r) { r[0] = nope; }"
    with pytest.raises(capi.MoptError) as e:
        capi.UserModel(bad, 2, 1, 1, 1)
    assert "nope" in str(e.value) and "user_model.cu(1)" in str(e.value)
    # f_df promised but not defined
    with pytest.raises(capi.MoptError):
        capi.UserModel(P2P_SRC, 6, 3, 3, 3, has_jacobian=True, set_size=12)


@pytest.mark.parametrize("args", [(0, 1, 1, 1), (17, 1, 1, 1), (2, 5, 1, 1), (2, 1, 0, 1), (2, 1, 3, 4), (2, 1, 4, 1)])
def test_user_model_shape_validation(args):
    from moptimizer_0_b200 import capi
    with pytest.raises(capi.MoptError):
        capi.UserModel(EXP_CURVE_SRC, *args)
    with pytest.raises(capi.MoptError):
        capi.UserModel(EXP_CURVE_SRC, 2, 1, 1, 1, set_size=25)
    with pytest.raises(capi.MoptError):
        capi.UserModel(EXP_CURVE_SRC, 2, 1, 1, 1, rot_offset=0)  # 3-vector does not fit in P = 2


# ------------------------------------------------------------------------------ GPU ----
@pytest.fixture(scope="module")
def ctx():
    from moptimizer_0_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def np_linearize(r, J, w=None, C=None):
    """computeHessian (linearization.h:143-154) in numpy: r (N, O), J (N, O, P)."""
    N, O, P = J.shape
    w = np.ones(N) if w is None else w
    C = np.eye(O) if C is None else C
    CJ = np.einsum("ok,nkp->nop", C, J)
    H = np.einsum("n,nop,noq->pq", w, J, CJ)
    b = np.einsum("n,nop,no->p", w, J, r @ C.T)
    return H, b, float(np.sum(r * r))


def curve_data(n, seed=5):
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 5.0, n)
    y = np.exp(0.3 * t + 0.1) + rng.normal(0.0, 0.2, n)
    return t, y


@pytest.mark.gpu
@pytest.mark.parametrize("store_dtype,compute_dtype", [(1, 1), (0, 1), (0, 0)])
@pytest.mark.parametrize("jac", [0, 1, 2])
def test_user_exp_curve_is_bit_identical_to_builtin(ctx, store_dtype, compute_dtype, jac):
    from moptimizer_0_b200 import capi
    um = capi.UserModel(EXP_CURVE_SRC, 2, 1, 1, 1, has_jacobian=True)
    n = 100_003  # ragged: not a multiple of the vector width or the CTA size
    t, y = curve_data(n)
    res = []
    for model in (capi.MODEL_EXP_CURVE, um.model):
        st = capi.Store(ctx, model, n, store_dtype)
        st.upload(0, t)
        st.upload(1, y)
        # the builtin's finite differences default to its series-form quotient (ExpCurveModel::finish_diff) with fp32
        # compute; a run-time compiled model has no such hook and gets the literal per-residual form, which the
        # builtin takes with MOPT_FLAG_GENERIC_KERNEL
        flags = capi.FLAG_GENERIC_KERNEL if (model == capi.MODEL_EXP_CURVE and jac != 0) else 0
        prob = capi.make_problem(model, jac, compute_dtype, loss=capi.LOSS_HUBER, loss_param=0.3, flags=flags)
        out = [ctx.linearize(st, prob, x) for x in ([0.0, 0.0], [0.29, 0.13])]
        cost = ctx.compute_cost(st, prob, [0.29, 0.13])
        lm = ctx.lm_minimize([st], [prob], [0.0, 0.0], max_iterations=30)
        res.append((out, cost, lm))
        st.close()
    (ob, cb, lb), (ou, cu, lu) = res
    for (Hb, bb, sb), (Hu, bu, su) in zip(ob, ou):
        assert np.array_equal(Hb, Hu) and np.array_equal(bb, bu) and sb == su
    assert cb == cu
    assert lb.status == lu.status and lb.sequence == lu.sequence and np.array_equal(lb.x, lu.x)
    assert np.array_equal(lb.trace, lu.trace)
    um.release()


@pytest.mark.gpu
def test_user_exp_curve_reference_known_answer(ctx):
    """tst/curve_fitting.cpp:110-117 through a user model: (0.291861, 0.131439) +- 5e-5 from (0, 0) in 15 iterations."""
    from moptimizer_0_b200 import capi
    um = capi.UserModel(EXP_CURVE_SRC, 2, 1, 1, 1, has_jacobian=True)
    t = np.array(FX["curve"]["t"], dtype=np.float64)
    y = np.array(FX["curve"]["y"], dtype=np.float64)
    st = capi.Store(ctx, um.model, len(t), capi.F64)
    st.upload(0, t)
    st.upload(1, y)
    for jac in (capi.JAC_ANALYTICAL, capi.JAC_FORWARD):
        prob = capi.make_problem(um.model, jac, capi.F64)
        r = ctx.lm_minimize([st], [prob], [0.0, 0.0], max_iterations=15, speculative=False)
        assert abs(r.x[0] - 0.291861) < 5e-5 and abs(r.x[1] - 0.131439) < 5e-5
    st.close()
    um.release()


@pytest.mark.gpu
@pytest.mark.parametrize("store_dtype,compute_dtype", [(1, 1), (0, 0)])
def test_user_p2p_with_setup_matches_builtin_and_oracle(ctx, store_dtype, compute_dtype):
    from moptimizer_0_b200 import capi
    from oracle import oracle_py as orc
    um = capi.UserModel(P2P_SRC, 6, 3, 3, 3, set_size=12, rot_offset=3)
    src, tgt, _, _ = fachada()
    n = src.shape[0]
    x = [1.0, 2.0, 3.0, 0.2, -0.3, 0.4]
    out = {}
    for model in (capi.MODEL_POINT2POINT, um.model):
        st = capi.Store(ctx, model, n, store_dtype)
        st.upload(0, src)
        st.upload(1, tgt)
        for jac in (capi.JAC_FORWARD, capi.JAC_CENTRAL):
            # consts[0] tells the user setup which Scalar the reference would run so3::Exp in (its eps guard)
            # the builtin model is held to the generic per-residual kernel (its default for finite differences is
            # the moment kernel), which is the code a user model is compiled into
            prob = capi.make_problem(model, jac, compute_dtype, consts=[1.0 if compute_dtype == 0 else 0.0],
                                     flags=capi.FLAG_GENERIC_KERNEL)
            out[(model, jac)] = ctx.linearize(st, prob, x)
        prob = capi.make_problem(model, capi.JAC_FORWARD, compute_dtype, consts=[1.0 if compute_dtype == 0 else 0.0],
                                 flags=capi.FLAG_GENERIC_KERNEL)
        out[(model, "lm")] = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50)
        if model != capi.MODEL_POINT2POINT:  # manifold update works for user models that name a rotation block
            probm = capi.make_problem(model, capi.JAC_FORWARD, compute_dtype, manifold=capi.MANIFOLD_SO3_LEFT,
                                      consts=[1.0 if compute_dtype == 0 else 0.0])
            out["lm_manifold"] = ctx.lm_minimize([st], [probm], [0.0] * 6, max_iterations=50)
        st.close()
    for jac in (capi.JAC_FORWARD, capi.JAC_CENTRAL):
        Hb, bb, sb = out[(capi.MODEL_POINT2POINT, jac)]
        Hu, bu, su = out[(um.model, jac)]
        assert np.array_equal(Hb, Hu) and np.array_equal(bb, bu) and sb == su
    lb, lu = out[(capi.MODEL_POINT2POINT, "lm")], out[(um.model, "lm")]
    assert lb.status == lu.status and lb.sequence == lu.sequence and np.array_equal(lb.x, lu.x)
    if compute_dtype == 1:
        oc = orc.Cost(orc.P2P, 6, 3, n, a=src, b=tgt, jac_mode=orc.JAC_FORWARD)
        Ho, bo, so = orc.linearize(oc, x)
        Hu, bu, su = out[(um.model, capi.JAC_FORWARD)]
        # finite differences amplify rounding by 1/h ~ 7e7: agreement to ~1e-7 is what fp64 allows
        assert rel_err(Hu, Ho) < 1e-6 and rel_err(bu, bo) < 1e-6 and abs(su - so) <= 1e-10 * so
        gt = FX["fachada"]["lm_numerical"]["x"] if "lm_numerical" in FX["fachada"] else None
        assert lu.status == "CONVERGED" and lu.executed_iterations == 5 and lu.sequence == "AAAAA"
        if gt is not None:
            assert np.max(np.abs(lu.x - np.array(gt))) < 1e-6
        lm = out["lm_manifold"]
        assert lm.status == "CONVERGED" and np.max(np.abs(lm.x - lu.x)) < 1e-6
    um.release()


def gauss_np(x, t, y):
    A, mu, sg = x
    d = t - mu
    g = np.exp(-(d * d) / (2 * sg * sg))
    r = (y - A * g)[:, None]
    J = np.stack([-g, -A * g * d / sg**2, -A * g * d * d / sg**3], axis=1)[:, None, :]
    return r, J


@pytest.mark.gpu
@pytest.mark.parametrize("store_dtype,compute_dtype,tol", [(1, 1, 1e-10), (0, 1, 1e-10), (0, 0, 1e-5)])
def test_user_gaussian_peak_against_numpy(ctx, store_dtype, compute_dtype, tol):
    from moptimizer_0_b200 import capi
    um = capi.UserModel(GAUSS_SRC, 3, 1, 1, 1, has_jacobian=True)
    rng = np.random.default_rng(11)
    n = 250_001
    t = rng.uniform(-4.0, 6.0, n)
    gt = np.array([2.5, 1.2, 0.8])
    y = gt[0] * np.exp(-((t - gt[1]) ** 2) / (2 * gt[2] ** 2)) + rng.normal(0.0, 0.05, n)
    if store_dtype == 0:
        t, y = t.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64)
    st = capi.Store(ctx, um.model, n, store_dtype)
    st.upload(0, t)
    st.upload(1, y)
    x = [2.0, 1.0, 1.0]
    r, J = gauss_np(x, t, y)
    # analytical, plain and with Huber weights + covariance scaling (tst/covariance.cpp:26-63 relation)
    Ho, bo, so = np_linearize(r, J)
    H, b, s = ctx.linearize(st, capi.make_problem(um.model, capi.JAC_ANALYTICAL, compute_dtype), x)
    assert rel_err(H, Ho) < tol and rel_err(b, bo) < tol and abs(s - so) <= tol * so
    k = 0.1
    e2 = np.sum(r * r, axis=1)
    w = np.where(e2 <= k * k, 1.0, k / np.sqrt(np.maximum(e2, 1e-300)))
    Hw, bw, _ = np_linearize(r, J, w=w, C=np.array([[0.5]]))
    H2, b2, s2 = ctx.linearize(st, capi.make_problem(um.model, capi.JAC_ANALYTICAL, compute_dtype, loss=capi.LOSS_HUBER,
                                                     loss_param=k, covariance=np.array([[0.5]])), x)
    assert rel_err(H2, Hw) < tol and rel_err(b2, bw) < tol and s2 == s
    # central differences agree with the analytical Jacobian (tst/differentiation.cpp:66-74 relation)
    if compute_dtype == 1:
        Hc, bc, _ = ctx.linearize(st, capi.make_problem(um.model, capi.JAC_CENTRAL, compute_dtype), x)
        assert rel_err(Hc, Ho) < 1e-6 and rel_err(bc, bo) < 1e-6
    # LM recovers the generating parameters
    res = ctx.lm_minimize([st], [capi.make_problem(um.model, capi.JAC_ANALYTICAL, compute_dtype)], x, max_iterations=50)
    assert np.max(np.abs(res.x - gt)) < 2e-3, res.x
    st.close()
    um.release()


@pytest.mark.gpu
@pytest.mark.parametrize("store_dtype,compute_dtype,tol", [(1, 1, 1e-10), (0, 0, 2e-5)])
def test_user_wide_polynomial_against_numpy(ctx, store_dtype, compute_dtype, tol):
    """P = 10, O = 2: 66 packed sums -> wide_pass_kernel, analytical and finite-difference, then the 10 x 10 solve."""
    from moptimizer_0_b200 import capi
    um = capi.UserModel(POLY_SRC, 10, 2, 2, 2, has_jacobian=True)
    rng = np.random.default_rng(3)
    n = 77_777
    t = rng.uniform(-1.0, 1.0, n)
    u = rng.uniform(0.5, 1.5, n)
    gt = rng.normal(0.0, 1.0, 10)
    V = np.vander(t, 10, increasing=True)
    p = V @ gt
    a = np.stack([t, u], axis=1)
    bdat = np.stack([p, u * p], axis=1) + rng.normal(0.0, 0.01, (n, 2))
    if store_dtype == 0:
        a, bdat = a.astype(np.float32).astype(np.float64), bdat.astype(np.float32).astype(np.float64)
        V = np.vander(a[:, 0], 10, increasing=True)
    st = capi.Store(ctx, um.model, n, store_dtype)
    st.upload(0, a)
    st.upload(1, bdat)
    x = rng.normal(0.0, 0.5, 10)
    px = V @ x
    r = np.stack([bdat[:, 0] - px, bdat[:, 1] - a[:, 1] * px], axis=1)
    J = np.stack([-V, -a[:, 1:2] * V], axis=1)
    C = np.array([[2.0, 0.3], [0.3, 1.0]])
    Ho, bo, so = np_linearize(r, J, C=C)
    H, b, s = ctx.linearize(st, capi.make_problem(um.model, capi.JAC_ANALYTICAL, compute_dtype, covariance=C), x)
    assert rel_err(H, Ho) < tol and rel_err(b, bo) < tol and abs(s - so) <= tol * so
    assert ctx.compute_cost(st, capi.make_problem(um.model, capi.JAC_ANALYTICAL, compute_dtype), x) == pytest.approx(so, rel=tol)
    if compute_dtype == 1:
        Hc, bc, sc = ctx.linearize(st, capi.make_problem(um.model, capi.JAC_CENTRAL, compute_dtype, covariance=C), x)
        assert rel_err(Hc, Ho) < 1e-6 and rel_err(bc, bo) < 1e-6 and abs(sc - s) <= 1e-12 * s
        res = ctx.lm_minimize([st], [capi.make_problem(um.model, capi.JAC_ANALYTICAL, compute_dtype)], x, max_iterations=30)
        ls = np.linalg.lstsq(np.concatenate([V, a[:, 1:2] * V]), np.concatenate([bdat[:, 0], bdat[:, 1]]), rcond=None)[0]
        assert np.max(np.abs(res.x - ls)) < 1e-6, (res.status, res.x - ls)
    st.close()
    um.release()


@pytest.mark.gpu
def test_user_model_errors_on_device(ctx):
    from moptimizer_0_b200 import capi
    um = capi.UserModel(P2P_SRC, 6, 3, 3, 3, set_size=12, rot_offset=3)  # no f_df
    st = capi.Store(ctx, um.model, 16, capi.F64)
    st.upload(0, np.zeros((16, 3)))
    st.upload(1, np.zeros((16, 3)))
    with pytest.raises(capi.MoptError) as e:
        ctx.linearize(st, capi.make_problem(um.model, capi.JAC_ANALYTICAL, capi.F64), [0.0] * 6)
    assert "f_df" in str(e.value)  # BaseModel::f_df throws for Jacobian-free models, model.h:66-70
    st.close()
    mid = um.model
    um.release()
    with pytest.raises(capi.MoptError):
        capi.Store(ctx, mid, 16, capi.F64)
