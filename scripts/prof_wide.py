"""Two launches of the n x n calibration pass (50 M observations, central differences, fp32) for ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi
from tests.common import camera_consts
ctx = capi.Context(0)
if len(sys.argv) > 1:
    ctx.set_launch(0, int(sys.argv[1]))
Cm = camera_consts()[12:]
x15 = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027, 600.0, 600.0, 320.0, 240.0, 0.05, -0.02, 0.001, -0.001, 0.005])
n = 50_000_000
st = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n, capi.F32)
st.generate(seed=3, gt=x15, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=Cm)
prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F32, consts=Cm)
for _ in range(2):
    ctx.linearize(st, prob, x15 * 0.999)
st.close(); ctx.close()
