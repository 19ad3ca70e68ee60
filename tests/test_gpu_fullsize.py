"""Full-size (-m gpu) checks through size-independent properties, at BASELINE.json's sizes where the oracle
would take minutes: additivity over shards, agreement of fp32 and fp64 compute, Jacobian-mode consistency,
and recovery of the generating parameters by the device-resident LM."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from tests.common import camera_consts, rel_err

pytestmark = pytest.mark.gpu
X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]


@pytest.fixture(scope="module")
def env():
    from moptimizer_0_b200 import capi
    c = capi.Context(0)
    yield capi, c
    c.close()


def test_p2p_100m_additivity_precision_and_oracle_sample(env):
    capi, ctx = env
    n = 100_000_000
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
    st.generate(seed=2, gt=X_GT, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    x = [0.1, -0.1, 0.05, 0.02, -0.01, 0.03]
    p32 = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.05)
    p64 = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F64, loss=capi.LOSS_HUBER, loss_param=0.05)
    H32, b32, s32 = ctx.linearize(st, p32, x)
    H64, b64, s64 = ctx.linearize(st, p64, x)
    # fp32 residual math + fp64 accumulation vs all-fp64 on the same fp32 inputs: the 1e-5 parity bar
    assert rel_err(H32, H64) < 1e-5 and rel_err(b32, b64) < 1e-5 and abs(s32 - s64) < 1e-5 * s64
    # additivity (tst/multiple_objectives.cpp as a property): two half-size stores of the same stream sum to the whole
    half = n // 2
    parts = []
    for first, cnt in ((0, half), (half, n - half)):
        sp = capi.Store(ctx, capi.MODEL_POINT2POINT, cnt, capi.F32)
        sp.generate(seed=2, gt=X_GT, first_index=first, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
        parts.append(ctx.linearize(sp, p64, x))
        sp.close()
    Hs, bs, ss = (parts[0][i] + parts[1][i] for i in range(3))
    assert rel_err(Hs, H64) < 1e-12 and rel_err(bs, b64) < 1e-12 and abs(ss - s64) < 1e-12 * s64
    # determinism: bit-identical run to run
    H32b, b32b, s32b = ctx.linearize(st, p32, x)
    assert np.array_equal(H32, H32b) and np.array_equal(b32, b32b) and s32 == s32b
    # oracle on the first 20 M rows (a fifth of the set) == device on a store holding exactly those rows
    m = 20_000_000
    src, tgt = st.download(0, np.float64, 0, m), st.download(1, np.float64, 0, m)
    sm = capi.Store(ctx, capi.MODEL_POINT2POINT, m, capi.F32)
    sm.upload(0, src)
    sm.upload(1, tgt)
    Hd, bd, sd = ctx.linearize(sm, p32, x)
    Ho, bo, so = orc.linearize(orc.Cost(orc.P2P, 6, 3, m, a=src, b=tgt, jac_mode=orc.JAC_ANALYTICAL,
                                        loss=orc.LOSS_HUBER, loss_param=0.05), x,
                              nthreads=max(2, orc.hardware_concurrency() or 8))
    assert rel_err(Hd, Ho) < 1e-5 and rel_err(bd, bo) < 1e-5 and abs(sd - so) < 1e-5 * so
    sm.close()
    # LM from x0 = 0 recovers the generating transform (noise sigma 0.01 over 1e8 points => ~1e-6)
    r = ctx.lm_minimize([st], [p32], [0.0] * 6, max_iterations=50)
    assert np.max(np.abs(r.x - np.array(X_GT))) < 2e-5, r.x
    st.close()


def test_curve_10m_central_difference(env):
    capi, ctx = env
    n = 10_000_000
    st = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
    st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
    x = [0.25, 0.15]
    ana = ctx.linearize(st, capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_ANALYTICAL, capi.F64), x)
    cen = ctx.linearize(st, capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F64), x)
    fwd = ctx.linearize(st, capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_FORWARD, capi.F64), x)
    assert rel_err(cen[0], ana[0]) < 1e-8 and rel_err(cen[1], ana[1]) < 1e-8   # central: O(h^2)
    assert rel_err(fwd[0], ana[0]) < 1e-6 and rel_err(fwd[1], ana[1]) < 1e-5   # forward: O(h)
    assert cen[2] == ana[2] == fwd[2]
    # oracle on a 1 M sample
    m = 1_000_000
    t, y = st.download(0, np.float64, 0, m)[:, 0], st.download(1, np.float64, 0, m)[:, 0]
    sm = capi.Store(ctx, capi.MODEL_EXP_CURVE, m, capi.F32)
    sm.upload(0, t)
    sm.upload(1, y)
    Hd, bd, sd = ctx.linearize(sm, capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F64), x)
    Ho, bo, so = orc.linearize(orc.Cost(orc.EXP_CURVE, 2, 1, m, a=t, b=y, jac_mode=orc.JAC_CENTRAL), x)
    assert rel_err(Hd, Ho) < 1e-7 and rel_err(bd, bo) < 1e-7 and abs(sd - so) < 1e-10 * so
    sm.close()
    r = ctx.lm_minimize([st], [capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F64)], [0.0, 0.0],
                        max_iterations=50)
    assert abs(r.x[0] - 0.3) < 1e-3 and abs(r.x[1] - 0.1) < 2e-3, r.x
    st.close()


def test_camera_50m_numerical(env):
    capi, ctx = env
    n = 50_000_000
    consts = camera_consts()
    x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027])
    T = orc.so3_convert6dof(x_gt)
    K, C = consts[:12].reshape(3, 4), consts[12:].reshape(4, 4)
    M = K @ T @ C
    st = capi.Store(ctx, capi.MODEL_PINHOLE, n, capi.F32)
    st.generate(seed=3, gt=M.reshape(-1), lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5)
    prob = capi.make_problem(capi.MODEL_PINHOLE, capi.JAC_CENTRAL, capi.F64, consts=consts)
    m = 500_000
    pts, pix = st.download(0, np.float64, 0, m), st.download(1, np.float64, 0, m)
    sm = capi.Store(ctx, capi.MODEL_PINHOLE, m, capi.F32)
    sm.upload(0, pts)
    sm.upload(1, pix)
    x = [0.0] * 6
    Hd, bd, sd = ctx.linearize(sm, prob, x)
    Ho, bo, so = orc.linearize(orc.Cost(orc.PINHOLE, 6, 2, m, a=pts, b=pix, consts=consts, jac_mode=orc.JAC_CENTRAL), x)
    assert rel_err(Hd, Ho) < 1e-6 and rel_err(bd, bo) < 1e-6 and abs(sd - so) < 1e-10 * so
    sm.close()
    r = ctx.lm_minimize([st], [prob], x, max_iterations=50)
    assert np.max(np.abs(r.x - x_gt)) < 1e-4, r.x
    st.close()


def test_camera_distort_50m_nxn_calibration(env):
    """BASELINE configs[4] at full size: pinhole + distortion, 50 M observations, 15 parameters, numerical Jacobian,
    15 x 15 device solve (wide kernel).  Size-independent properties: additivity over two stores, fp32 compute
    against fp64 compute on the same data, central against forward differences; an oracle check on a sample; and
    recovery of the generating parameters by the device LM."""
    capi, ctx = env
    n = 50_000_000
    Cm = camera_consts()[12:]
    x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027, 600.0, 600.0, 320.0, 240.0, 0.05, -0.02, 0.001,
                     -0.001, 0.005])
    st = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n, capi.F32)
    st.generate(seed=3, gt=x_gt, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=Cm)
    x = x_gt * (1.0 + 0.002 * np.cos(np.arange(15)))
    p64 = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F64, consts=Cm)
    p32 = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F32, consts=Cm)
    H, b, s = ctx.linearize(st, p64, x)
    d = np.sqrt(np.diag(H))
    scale = np.outer(d, d)
    # additivity: two half-size stores filled with the two halves of the same set (generator indexed globally)
    halves = []
    for first in (0, n // 2):
        h = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n // 2, capi.F32)
        h.generate(seed=3, gt=x_gt, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=Cm,
                   first_index=first)
        halves.append(ctx.linearize(h, p64, x))
        h.close()
    Hs, bs, ss = (halves[0][k] + halves[1][k] for k in range(3))
    assert np.max(np.abs(Hs - H) / scale) < 1e-11 and abs(ss - s) <= 1e-12 * s
    assert np.max(np.abs(bs - b) / (d * np.sqrt(s))) < 1e-11
    # fp32 residual math (fp64 accumulation) against fp64: the sum agrees to fp32 rounding.  The float finite-difference
    # Jacobian carries eps_f32 * |pixel| / h_j of noise per residual with the reference's step h_j = sqrt(eps_f32) |x_j|
    # (linearization.h:85-87) — tens of percent for the 1e-3-sized distortion parameters, which biases diag(H) by
    # noise^2 — so only the well-scaled block (extrinsics, focal lengths, principal point) is compared, loosely.
    H32, b32, s32 = ctx.linearize(st, p32, x)
    assert abs(s32 - s) <= 1e-5 * s
    assert np.max(np.abs(H32 - H)[:10, :10] / scale[:10, :10]) < 5e-2
    # ... which is why the fp32 throughput path forms the same quotients over a common denominator / from the
    # parameter-wise affine structure instead (wide_pass_kernel AFFINE_FD): the whole 15 x 15 system then agrees to
    # fp32 rounding; the per-residual form stays selectable (MOPT_FLAG_GENERIC_KERNEL) and keeps its noise floor.
    e32 = (np.max(np.abs(H32 - H) / scale), np.max(np.abs(b32 - b) / (d * np.sqrt(s))))
    print("camera15 50M fp32 vs fp64 compute: H %.2e b %.2e (scaled)" % e32)
    assert e32[0] < 2e-5 and e32[1] < 2e-4, e32
    Hg, _, sg = ctx.linearize(st, capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F32, consts=Cm,
                                                    flags=capi.FLAG_GENERIC_KERNEL), x)
    assert abs(sg - s32) <= 1e-6 * s32 and np.max(np.abs(Hg - H)[:10, :10] / scale[:10, :10]) < 5e-2
    # forward differences against central ones (fp64): first-order truncation only
    Hf, bf, sf = ctx.linearize(st, capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_FORWARD, capi.F64, consts=Cm), x)
    assert sf == s and np.max(np.abs(Hf - H) / scale) < 1e-4
    # oracle on the first 200 k observations
    m = 200_000
    pts, pix = st.download(0, np.float64, 0, m), st.download(1, np.float64, 0, m)
    sm = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, m, capi.F32)
    sm.upload(0, pts)
    sm.upload(1, pix)
    Hd, bd, sd = ctx.linearize(sm, p64, x)
    Ho, bo, so = orc.linearize(orc.Cost(orc.PINHOLE_DISTORT, 15, 2, m, a=pts, b=pix, consts=Cm, jac_mode=orc.JAC_CENTRAL), x)
    do = np.sqrt(np.diag(Ho))
    assert abs(sd - so) <= 1e-10 * so
    assert np.max(np.abs(Hd - Ho) / np.outer(do, do)) < 1e-5 and np.max(np.abs(bd - bo) / (do * np.sqrt(so))) < 1e-5
    sm.close()
    # the device LM recovers the generating parameters from a perturbed start
    x0 = x_gt.copy()
    x0[:6] = 0.0
    x0[6:10] *= 1.02
    x0[10:] = 0.0
    r = ctx.lm_minimize([st], [p64], x0, max_iterations=50)
    assert np.max(np.abs(r.x[:6] - x_gt[:6])) < 2e-4, r.x
    assert np.max(np.abs(r.x[6:10] / x_gt[6:10] - 1.0)) < 1e-4, r.x
    st.close()


def test_p2p_20m_finite_differences_moment_vs_per_residual(env):
    """Point2point finite differences run on the moment kernel (the difference quotient of an affine residual is affine
    in the source point); MOPT_FLAG_GENERIC_KERNEL forms the quotient per residual like linearization.h:97-111.  In
    fp64 the two differ only by that subtraction's rounding (eps / h ~ 1e-8 per entry), and both sit at the
    truncation distance from the analytical Jacobian."""
    capi, ctx = env
    n = 20_000_000
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
    st.generate(seed=2, gt=X_GT, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    x = [0.45, -0.25, 0.22, 0.09, -0.06, 0.07]
    kw = dict(loss=capi.LOSS_HUBER, loss_param=0.05)
    Ha, ba, sa = ctx.linearize(st, capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F64, **kw), x)
    for jac, trunc in ((capi.JAC_FORWARD, 1e-6), (capi.JAC_CENTRAL, 1e-7)):  # central: the eps / h rounding floor
        Hm, bm, sm = ctx.linearize(st, capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64, **kw), x)
        Hg, bg, sg = ctx.linearize(st, capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64,
                                                         flags=capi.FLAG_GENERIC_KERNEL, **kw), x)
        assert abs(sm - sg) <= 1e-12 * sg and abs(sm - sa) <= 1e-12 * sa
        assert rel_err(Hm, Hg) < 1e-7 and rel_err(bm, bg) < 1e-7, (rel_err(Hm, Hg), rel_err(bm, bg))
        assert rel_err(Hm, Ha) < trunc and rel_err(bm, ba) < trunc, (jac, rel_err(Hm, Ha), rel_err(bm, ba))
    st.close()
