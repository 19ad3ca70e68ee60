"""Device LM on the reference's own small case (tst/point2point: fachada cloud, 29 310 points) for a launch list."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi
from tests.common import fachada

ctx = capi.Context(0)
src, tgt, _, _ = fachada()
st = capi.Store(ctx, capi.MODEL_POINT2POINT, src.shape[0], capi.F64)
st.upload(0, src)
st.upload(1, tgt)
for jac in (capi.JAC_ANALYTICAL, capi.JAC_FORWARD):
    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, capi.F64)
    for spec in (True, False):
        ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, speculative=spec)
        best = 1e9
        for _ in range(10):
            t0 = time.perf_counter()
            r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, speculative=spec)
            best = min(best, time.perf_counter() - t0)
        print(f"jac={jac} speculative={spec}: {r.status} {r.executed_iterations} iterations {r.num_passes} passes, best {best*1e6:.1f} us "
              f"-> {r.executed_iterations / best:.0f} LM iterations/s", flush=True)
st.close()
ctx.close()
