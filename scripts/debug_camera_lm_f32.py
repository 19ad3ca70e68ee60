"""Diagnostic: LM on the camera model with fp32 compute at several sizes, both finite-difference forms."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi
from oracle import oracle_py as orc
from tests.common import camera_consts

ctx = capi.Context(0)
consts = camera_consts()
x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027])
M = consts[:12].reshape(3, 4) @ orc.so3_convert6dof(x_gt) @ consts[12:].reshape(4, 4)
np.set_printoptions(linewidth=200, precision=6)
for n in (200_000, 5_000_000, 50_000_000):
    st = capi.Store(ctx, capi.MODEL_PINHOLE, n, capi.F32)
    st.generate(seed=3, gt=M.reshape(-1), lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5)
    for cd, flags, name in ((capi.F64, 0, "f64"), (capi.F32, 0, "f32 common-denominator"), (capi.F32, 1, "f32 per-residual")):
        for spec in (True, False):
            prob = capi.make_problem(capi.MODEL_PINHOLE, capi.JAC_CENTRAL, cd, consts=consts, flags=flags)
            r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, speculative=spec)
            print(f"n={n} {name} speculative={spec}: {r.status} it={r.executed_iterations} passes={r.num_passes} "
                  f"seq={r.sequence} x_err={np.max(np.abs(r.x - x_gt)):.3e} x={r.x}")
            if n == 50_000_000 and cd == capi.F32 and flags == 0 and spec:
                for t in r.trace:
                    print("   it=%d k=%d y0=%.10e yi=%.10e rho=%.4e lambda=%.4e nu=%g acc=%d" % tuple(t))
    st.close()
ctx.close()
