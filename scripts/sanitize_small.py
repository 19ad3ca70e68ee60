"""Every kernel of libmopt_b200.so once, at small ragged sizes, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
    compute-sanitizer --tool synccheck python scripts/sanitize_small.py

No torch, no oracle: the point is the memory / shared-memory / barrier behaviour of the kernels, the
numbers are checked elsewhere (tests/).  Prints one line per case so a sanitizer report can be placed.
(compute-sanitizer is closed on this round's GPU pool — the run answered "closed on this pool" — so only the plain
run exists: every kernel variant at a ragged size completes without a CUDA error, gpurun_out/sanitize_plain.log.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi  # noqa: E402

ctx = capi.Context(0)
rng = np.random.default_rng(0)
N = int(os.environ.get("MOPT_SANITIZE_N", "4099"))  # not a multiple of 4 / 32 / the CTA size: ragged tails everywhere


def say(name, *vals):
    print(name, *("%.6g" % v for v in vals), flush=True)


# ---- point2point: moment kernel (analytical, finite differences), generic kernel, cost, masked --------------
src = rng.uniform(0, 10, (N, 3))
tgt = src + rng.normal(0, 0.01, (N, 3)) + np.array([0.1, -0.2, 0.05])
for sd, cd in ((capi.F32, capi.F32), (capi.F32, capi.F64), (capi.F64, capi.F64)):
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, N, sd)
    st.upload(0, src)
    st.upload(1, tgt)
    for jac in (capi.JAC_ANALYTICAL, capi.JAC_FORWARD, capi.JAC_CENTRAL):
        for flags in (0, capi.FLAG_GENERIC_KERNEL):
            if jac == capi.JAC_ANALYTICAL and flags:
                continue
            for loss, lp in ((capi.LOSS_NONE, 0.0), (capi.LOSS_HUBER, 0.05), (capi.LOSS_GEMAN_MCCLURE, 1.0)):
                prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, cd, loss=loss, loss_param=lp, flags=flags)
                H, b, s = ctx.linearize(st, prob, [0.01, 0.02, 0.03, 0.01, -0.02, 0.03])
                c = ctx.compute_cost(st, prob, [0.01, 0.02, 0.03, 0.01, -0.02, 0.03])
                say(f"p2p store={sd} compute={cd} jac={jac} flags={flags} loss={loss}", H[0, 0], b[0], s, c)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, cd, covariance=np.diag([2.0, 1.0, 0.5]))
    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=8)
    say(f"p2p LM store={sd} compute={cd} {r.status}", r.executed_iterations, r.final_cost)
    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=4, speculative=False)
    say(f"p2p LM (reference pass order) {r.status}", r.executed_iterations)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, cd, variant=capi.P2P_LEFT,
                             manifold=capi.MANIFOLD_SO3_LEFT)
    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=4)
    say(f"p2p LM manifold {r.status}", r.executed_iterations)
    st.close()

# ---- model->update(x): grid index + re-association + masked passes -----------------------------------------
st = capi.Store(ctx, capi.MODEL_POINT2POINT, N, capi.F32)
st.upload(0, src)
ix = capi.NNIndex(ctx, tgt[: N - 7], 0.5, dtype=capi.F32)
capi.store_set_target(st, ix)
m = capi.store_reassociate(st, [0.0] * 6)
for jac in (capi.JAC_ANALYTICAL, capi.JAC_CENTRAL):
    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.05)
    H, b, s = ctx.linearize(st, prob, [0.0] * 6)
    say(f"icp matched={m} jac={jac}", H[0, 0], s)
r = ctx.lm_minimize([st], [capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32)], [0.0] * 6,
                    max_iterations=5)
say(f"icp LM {r.status}", r.executed_iterations)
capi.store_set_target(st, None)
ix.close()
st.close()

# ---- point distance (cost only), curve, Michaelis-Menten, Powell -------------------------------------------
st = capi.Store(ctx, capi.MODEL_POINT_DIST, N, capi.F64)
st.upload(0, src)
st.upload(1, tgt)
say("point_dist cost", ctx.compute_cost(st, capi.make_problem(capi.MODEL_POINT_DIST, capi.JAC_FORWARD, capi.F64), []))
st.close()
t = np.linspace(0, 5, N)
y = np.exp(0.3 * t + 0.1) + rng.normal(0, 0.2, N)
for model in (capi.MODEL_EXP_CURVE, capi.MODEL_MICHAELIS_MENTEN):
    for sd, cd in ((capi.F32, capi.F32), (capi.F32, capi.F64), (capi.F64, capi.F64)):
        st = capi.Store(ctx, model, N, sd)
        st.upload(0, t)
        st.upload(1, y)
        for jac in (capi.JAC_ANALYTICAL, capi.JAC_FORWARD, capi.JAC_CENTRAL):
            H, b, s = ctx.linearize(st, capi.make_problem(model, jac, cd), [0.25, 0.15])
            say(f"model={model} store={sd} compute={cd} jac={jac}", H[0, 0], b[0], s)
        r = ctx.lm_minimize([st], [capi.make_problem(model, capi.JAC_CENTRAL, cd)], [0.2, 0.2], max_iterations=5)
        say(f"model={model} LM {r.status}", r.executed_iterations)
        st.close()
st = capi.Store(ctx, capi.MODEL_POWELL, 1, capi.F64)
for jac in (capi.JAC_ANALYTICAL, capi.JAC_FORWARD):
    r = ctx.lm_minimize([st], [capi.make_problem(capi.MODEL_POWELL, jac, capi.F64)], [3.0, -1.0, 0.0, 4.0], max_iterations=10)
    say(f"powell jac={jac} {r.status}", r.executed_iterations, r.final_cost)
st.close()

# ---- cameras: dense kernel (P = 6) and wide kernel (P = 15), both finite-difference forms -------------------
K = np.array([586.0, 0, 638.0, 0, 0, 722.0, 323.0, 0, 0, 0, 1, 0])
Cm = np.eye(4)
consts6 = np.concatenate([K, Cm.reshape(-1)])
pts = np.column_stack([rng.uniform(-1, 1, N), rng.uniform(-0.5, 0.5, N), rng.uniform(2, 5, N)])
pix = np.column_stack([586.0 * pts[:, 0] / pts[:, 2] + 638.0, 722.0 * pts[:, 1] / pts[:, 2] + 323.0]) + rng.normal(0, 0.5, (N, 2))
x15 = np.array([0.01, -0.02, 0.03, 0.01, 0.02, -0.01, 586.0, 722.0, 638.0, 323.0, -0.1, 0.05, 0.001, -0.0007, 0.01])
for model, consts, x in ((capi.MODEL_PINHOLE, consts6, x15[:6]), (capi.MODEL_PINHOLE_DISTORT, Cm.reshape(-1), x15)):
    for sd, cd in ((capi.F32, capi.F32), (capi.F32, capi.F64), (capi.F64, capi.F64)):
        st = capi.Store(ctx, model, N, sd)
        st.upload(0, pts)
        st.upload(1, pix)
        for jac in (capi.JAC_FORWARD, capi.JAC_CENTRAL):
            for flags in (0, capi.FLAG_GENERIC_KERNEL):
                prob = capi.make_problem(model, jac, cd, consts=consts, flags=flags, loss=capi.LOSS_HUBER, loss_param=2.0,
                                         covariance=np.array([[1.5, 0.1], [0.1, 0.8]]))
                H, b, s = ctx.linearize(st, prob, x)
                c = ctx.compute_cost(st, prob, x)
                say(f"camera model={model} store={sd} compute={cd} jac={jac} flags={flags}", H[0, 0], b[0], s, c)
        r = ctx.lm_minimize([st], [capi.make_problem(model, capi.JAC_CENTRAL, cd, consts=consts)], x, max_iterations=4)
        say(f"camera model={model} LM {r.status}", r.executed_iterations)
        st.close()

# ---- device generator, download, ingest round trip ----------------------------------------------------------
st = capi.Store(ctx, capi.MODEL_POINT2POINT, N, capi.F32)
st.generate(seed=2, gt=[0.5, -0.3, 0.2, 0.1, -0.05, 0.08], lo=(0, 0, 0), hi=(10, 10, 10), noise_sigma=0.01,
            outlier_fraction=0.05, outlier_range=1.0)
a = st.download(0, np.float32)
say("generate/download", float(a[0, 0]), float(a[-1, 2]))
st.close()

# ---- a run-time compiled user model (dense and wide kernels through NVRTC) ----------------------------------
SRC = """
template <typename T> __device__ void mopt_f(const T* s, const T* a, const T* b, T* r) { r[0] = b[0] - exp(fma(s[0], a[0], s[1])); }
template <typename T> __device__ void mopt_f_df(const T* s, const T* a, const T* b, T* r, T* J) {
  const T ex = exp(fma(s[0], a[0], s[1])); r[0] = b[0] - ex; J[0] = -a[0] * ex; J[1] = -ex; }
"""
try:
    um = capi.UserModel(SRC, 2, 1, 1, 1, has_jacobian=True)
    st = capi.Store(ctx, um.model, N, capi.F32)
    st.upload(0, t)
    st.upload(1, y)
    for jac in (capi.JAC_ANALYTICAL, capi.JAC_CENTRAL):
        H, b, s = ctx.linearize(st, capi.make_problem(um.model, jac, capi.F32), [0.25, 0.15])
        say(f"user model jac={jac}", H[0, 0], b[0], s)
    st.close()
except Exception as e:  # NVRTC missing on the box: not what this script is about
    print("user model skipped:", e)

ctx.close()
print("sanitize_small: done")
