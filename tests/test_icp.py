"""Correspondence re-association, `model->update(x)`: declared by the reference
(include/moptimizer/model.h:24-26, called at src/levenberg_marquadt_dyn.cpp:54) and never implemented there.
Oracle = brute-force exact nearest neighbour (oracle::Point2PointICP), itself checked against scipy's cKDTree;
device = uniform-grid search (csrc/mopt_icp.cu) checked against the oracle, stand-alone and inside the LM loop."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from tests.common import fachada, rel_err

X_TRUE = np.array([0.12, -0.08, 0.05, 0.03, -0.02, 0.04])
MAX_DIST = 0.6


def clouds(n_src=3000, n_tgt=5000, seed=0, outliers=200):
    """target = a fachada subset; source = T_true^-1 (other subset of the target) + noise, plus far outliers."""
    src_all, _, _, _ = fachada()
    rng = np.random.default_rng(seed)
    tgt = src_all[rng.choice(src_all.shape[0], n_tgt, replace=False)]
    T = orc.so3_convert6dof(X_TRUE)
    pick = tgt[rng.choice(n_tgt, n_src - outliers, replace=False)] + rng.normal(0, 0.002, (n_src - outliers, 3))
    src = (pick - T[:3, 3]) @ T[:3, :3]          # R^T (q - t)
    far = rng.uniform(-50, -40, (outliers, 3))   # nothing within MAX_DIST of these
    return np.ascontiguousarray(np.vstack([src, far])), np.ascontiguousarray(tgt)


def icp_cost(src, tgt, jac=orc.JAC_ANALYTICAL, variant=orc.P2P_EXACT, update_x=None, **kw):
    return orc.Cost(orc.P2P_ICP, 6, 3, src.shape[0], a=src, jac_mode=jac, variant=variant, target=tgt,
                    max_dist=MAX_DIST, update_x=update_x, **kw)


def brute_nn(src, tgt, x, max_dist):
    from scipy.spatial import cKDTree
    T = orc.so3_convert6dof(x)
    q = src @ T[:3, :3].T + T[:3, 3]
    d, idx = cKDTree(tgt).query(q, k=1, distance_upper_bound=max_dist * (1 + 1e-12))
    ok = np.isfinite(d) & (d <= max_dist)
    return ok, np.where(ok, idx, 0)


def test_oracle_update_matches_kdtree():
    src, tgt = clouds()
    for x in ([0.0] * 6, X_TRUE, [0.3, 0.1, -0.2, 0.05, 0.05, -0.05]):
        ok, idx = brute_nn(src, tgt, x, MAX_DIST)
        # the oracle exposes its correspondences through the linearization: compare with a plain p2p oracle
        # cost built from the kd-tree's matches
        H, b, s = orc.linearize(icp_cost(src, tgt), x)
        ref = orc.Cost(orc.P2P, 6, 3, int(ok.sum()), a=src[ok], b=tgt[idx[ok]], jac_mode=orc.JAC_ANALYTICAL)
        Hr, br, sr = orc.linearize(ref, x)
        assert rel_err(H, Hr) < 1e-12 and rel_err(b, br) < 1e-12 and s == pytest.approx(sr, rel=1e-12)
        assert ok.sum() < src.shape[0]  # the far outliers are unmatched


def test_oracle_icp_lm_recovers_the_motion():
    src, tgt = clouds()
    r = orc.lm_minimize([icp_cost(src, tgt, cost_threads=4)], [0.0] * 6, max_iterations=30)
    assert np.allclose(r.x, X_TRUE, atol=2e-3), r.x


# ------------------------------------------------------------------------------------------ device ----
@pytest.fixture(scope="module")
def env():
    from moptimizer_0_b200 import capi
    c = capi.Context(0)
    yield capi, c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [0, 1])
def test_device_reassociation_matches_brute_force(env, dtype):
    capi, ctx = env
    src, tgt = clouds()
    n = src.shape[0]
    if dtype == 0:
        src, tgt = src.astype(np.float32).astype(np.float64), tgt.astype(np.float32).astype(np.float64)
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, dtype)
    st.upload(0, src)
    ix = capi.NNIndex(ctx, tgt, MAX_DIST, dtype=capi.F64)
    capi.store_set_target(st, ix)
    for x in ([0.0] * 6, X_TRUE, [0.3, 0.1, -0.2, 0.05, 0.05, -0.05]):
        matched = capi.store_reassociate(st, x)
        ok, idx = brute_nn(src, tgt, x, MAX_DIST)
        assert matched == int(ok.sum())
        got = st.download(1)
        assert np.array_equal(np.isnan(got[:, 0]), ~ok)
        assert np.array_equal(got[ok], tgt[idx[ok]])          # the same target points, bit for bit
        # passes skip the unmatched residuals exactly like `f` returning false (linearization.h:102,144)
        for jac, tol in ((capi.JAC_ANALYTICAL, 1e-10), (capi.JAC_FORWARD, 1e-6)):
            prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64, loss=capi.LOSS_HUBER, loss_param=0.05)
            H, b, s = ctx.linearize(st, prob, x)
            Ho, bo, so = orc.linearize(icp_cost(src, tgt, jac, loss=orc.LOSS_HUBER, loss_param=0.05), x)
            assert rel_err(H, Ho) < tol and rel_err(b, bo) < tol and abs(s - so) <= 1e-10 * so
            assert ctx.compute_cost(st, prob, x) == pytest.approx(so, rel=1e-10)
    # fp32 compute path with masked residuals
    prob32 = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, dtype and capi.F64 or capi.F32)
    H, b, s = ctx.linearize(st, prob32, x)
    Ho, bo, so = orc.linearize(icp_cost(src, tgt), x)
    assert rel_err(H, Ho) < 1e-5 and rel_err(b, bo) < 1e-5 and abs(s - so) <= 1e-5 * so
    capi.store_set_target(st, None)
    ix.close()
    st.close()


@pytest.mark.gpu
@pytest.mark.parametrize("jac,variant", [(0, 0), (1, 0)])
def test_device_icp_lm_matches_oracle(env, jac, variant):
    capi, ctx = env
    src, tgt = clouds()
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F64)
    st.upload(0, src)
    ix = capi.NNIndex(ctx, tgt, MAX_DIST, dtype=capi.F64)
    capi.store_set_target(st, ix)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64, variant=variant)
    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=30)
    ro = orc.lm_minimize([icp_cost(src, tgt, jac, variant, cost_threads=4)], [0.0] * 6, max_iterations=30)
    assert np.allclose(r.x, X_TRUE, atol=2e-3)
    assert np.allclose(r.x, ro.x, atol=1e-6)
    k = min(len(r.sequence), len(ro.sequence))
    decisive = [i for i in range(k) if abs(ro.trace[i, 2] - ro.trace[i, 3]) > 1e-9 * ro.trace[i, 2]]
    cut = (decisive[-1] + 1) if decisive else 0
    assert cut >= 3 and r.sequence[:cut] == ro.sequence[:cut]
    # y0 after each re-association: the analytical path reproduces the oracle's iterates to rounding; with
    # forward differences the iterates differ by the finite-difference noise (1/h ~ 7e7 times eps)
    assert np.allclose(r.trace[:cut, 2], ro.trace[:cut, 2], rtol=1e-9 if jac == 0 else 1e-4)
    ix.close()
    st.close()


@pytest.mark.gpu
def test_device_reassociation_large_cloud_throughput(env):
    """1 M x 1 M points: property checks (every match within the radius, idempotence) and a rate print-out."""
    import time
    capi, ctx = env
    rng = np.random.default_rng(5)
    tgt = rng.uniform(0, 20, (1_000_000, 3)).astype(np.float32)
    src = tgt[rng.permutation(1_000_000)] + rng.normal(0, 0.01, (1_000_000, 3)).astype(np.float32)
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F32)
    st.upload(0, src)
    ix = capi.NNIndex(ctx, tgt, 0.1, dtype=capi.F32)
    capi.store_set_target(st, ix)
    x = [0.0] * 6
    capi.store_reassociate(st, x)
    t0 = time.perf_counter()
    matched = capi.store_reassociate(st, x)
    dt = time.perf_counter() - t0
    got = st.download(1, np.float32)
    ok = ~np.isnan(got[:, 0])
    assert matched == int(ok.sum()) and matched > 990_000
    assert np.all(np.linalg.norm(got[ok].astype(np.float64) - src[ok], axis=1) <= 0.1 + 1e-6)
    again = capi.store_reassociate(st, x)
    assert again == matched and np.array_equal(st.download(1, np.float32)[ok], got[ok])
    print(f"re-association 1M x 1M: {dt * 1e3:.2f} ms ({1.0 / dt:.1f} M queries/s)")
    ix.close()
    st.close()
