"""SO(3) manifold update: the opt-in that finishes the reference's "TODO Manifold operation"
(src/levenberg_marquadt_dyn.cpp:82) with so3::Exp / so3::Log (src/so3.cpp:43-57,96-105).  The default stays
additive; these tests cover the opt-in on the oracle (CPU) and its parity on the device (gpu)."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from tests.common import fachada, rel_err

X = [1.0, 2.0, 3.0, 0.2, -0.3, 0.4]


def ocost(src, tgt, jac, variant=orc.P2P_LEFT, manifold=orc.MANIFOLD_SO3_LEFT, **kw):
    return orc.Cost(orc.P2P, 6, 3, src.shape[0], a=src, b=tgt, jac_mode=jac, variant=variant, manifold=manifold, **kw)


def test_oracle_left_jacobian_matches_manifold_finite_differences():
    src, tgt, _, _ = fachada()
    Ha, ba, sa = orc.linearize(ocost(src, tgt, orc.JAC_ANALYTICAL), X)
    Hc, bc, sc = orc.linearize(ocost(src, tgt, orc.JAC_CENTRAL), X)
    assert rel_err(Ha, Hc) < 1e-6 and rel_err(ba, bc) < 1e-6 and sa == pytest.approx(sc, rel=1e-14)
    # at omega = 0 the left-perturbation Jacobian coincides with the additive one
    H0, b0, _ = orc.linearize(ocost(src, tgt, orc.JAC_ANALYTICAL), [0.0] * 6)
    He, be, _ = orc.linearize(ocost(src, tgt, orc.JAC_ANALYTICAL, variant=orc.P2P_EXACT, manifold=0), [0.0] * 6)
    assert rel_err(H0, He) < 1e-14 and rel_err(b0, be) < 1e-14


def test_oracle_manifold_lm_reaches_the_same_transform():
    src, tgt, R, t = fachada()
    add = orc.lm_minimize([ocost(src, tgt, orc.JAC_ANALYTICAL, variant=orc.P2P_EXACT, manifold=0, cost_threads=4)],
                          [0.0] * 6, max_iterations=50)
    man = orc.lm_minimize([ocost(src, tgt, orc.JAC_ANALYTICAL, cost_threads=4)], [0.0] * 6, max_iterations=50)
    assert add.status == man.status == "CONVERGED"
    Tm = orc.so3_convert6dof(man.x)
    assert np.allclose(Tm[:3, :3], R, atol=1e-9) and np.allclose(Tm[:3, 3], t, atol=1e-8)
    assert np.allclose(man.x, add.x, atol=1e-8)
    num = orc.lm_minimize([ocost(src, tgt, orc.JAC_FORWARD, cost_threads=4)], [0.0] * 6, max_iterations=50)
    assert num.status == "CONVERGED" and np.allclose(num.x, man.x, atol=1e-8)


# ------------------------------------------------------------------------------------------ device ----
@pytest.fixture(scope="module")
def env():
    from moptimizer_0_b200 import capi
    c = capi.Context(0)
    yield capi, c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("jac,tol", [(0, 1e-10), (1, 1e-6), (2, 1e-6)])
def test_device_manifold_linearization_matches_oracle(env, jac, tol):
    capi, ctx = env
    src, tgt, _, _ = fachada()
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F64)
    st.upload(0, src)
    st.upload(1, tgt)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64, variant=capi.P2P_LEFT,
                             manifold=capi.MANIFOLD_SO3_LEFT)
    H, b, s = ctx.linearize(st, prob, X)
    Ho, bo, so = orc.linearize(ocost(src, tgt, jac), X)
    assert rel_err(H, Ho) < tol and rel_err(b, bo) < tol and abs(s - so) <= 1e-10 * so
    st.close()


@pytest.mark.gpu
@pytest.mark.parametrize("jac", [0, 1])
@pytest.mark.parametrize("speculative", [True, False])
def test_device_manifold_lm_matches_oracle(env, jac, speculative):
    capi, ctx = env
    src, tgt, R, t = fachada()
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F64)
    st.upload(0, src)
    st.upload(1, tgt)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64, variant=capi.P2P_LEFT,
                             manifold=capi.MANIFOLD_SO3_LEFT)
    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, speculative=speculative)
    ro = orc.lm_minimize([ocost(src, tgt, jac, cost_threads=4)], [0.0] * 6, max_iterations=50)
    assert r.status == ro.status == "CONVERGED"
    assert r.executed_iterations == ro.executed_iterations and r.sequence == ro.sequence
    assert np.allclose(r.x, ro.x, atol=1e-6)
    T = orc.so3_convert6dof(r.x)
    assert np.allclose(T[:3, :3], R, atol=1e-9) and np.allclose(T[:3, 3], t, atol=1e-8)
    st.close()


@pytest.mark.gpu
def test_manifold_is_rejected_for_models_without_rotation(env):
    capi, ctx = env
    st = capi.Store(ctx, capi.MODEL_EXP_CURVE, 4, capi.F64)
    with pytest.raises(capi.MoptError, match="rotation vector"):
        ctx.linearize(st, capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_FORWARD, capi.F64, manifold=1), [0.0, 0.0])
    st.close()
