"""Small-problem LM: the persistent one-launch kernel against the launch-per-trial path (MOPT_LM_MONO=0), on
tst/point2point's cloud (29 310 points, fp64) and on synthetic sets of growing size (fp32, Huber)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi
from tests.common import fachada

ctx = capi.Context(0)
def rate(st, prob, reps=30, **kw):
    ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, **kw)
    ts = []
    for _ in range(reps):
        ctx.synchronize(); t0 = time.perf_counter()
        r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, **kw)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return r, dt
src, tgt, _, _ = fachada()
st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F64)
st.upload(0, src); st.upload(1, tgt)
for jac, name in ((capi.JAC_ANALYTICAL, "analytical"), (capi.JAC_FORWARD, "forward differences")):
    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64, variant=capi.P2P_EXACT)
    r, dt = rate(st, prob)
    print(f"fachada fp64 {name}: {r.status} {r.sequence} {r.executed_iterations} iterations {r.num_passes} passes, "
          f"{dt * 1e6:.1f} us per solve = {r.executed_iterations / dt:.0f} LM iterations/s, {dt / r.num_passes * 1e6:.1f} us per pass; x = {np.round(r.x, 8)}", flush=True)
st.close()
X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]
for n in (20_000, 100_000, 250_000, 1_000_000, 4_000_000):
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
    st.generate(seed=2, gt=X_GT, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.05)
    r, dt = rate(st, prob, reps=10)
    print(f"synthetic fp32 n = {n}: {r.status} {r.sequence} {r.executed_iterations} it {r.num_passes} passes, {dt * 1e6:.1f} us, "
          f"{dt / r.num_passes * 1e6:.1f} us per pass", flush=True)
    st.close()
