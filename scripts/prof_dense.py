"""A few launches of the finite-difference pass kernels for ncu (camera 50 M P=6 dense, P=15 wide, curve 10 M)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi
from oracle import oracle_py as orc
from tests.common import camera_consts

ctx = capi.Context(0)
if len(sys.argv) > 2:
    ctx.set_launch(int(sys.argv[1]), int(sys.argv[2]))
consts = camera_consts()
x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027])
M = consts[:12].reshape(3, 4) @ orc.so3_convert6dof(x_gt) @ consts[12:].reshape(4, 4)
n = 50_000_000
st = capi.Store(ctx, capi.MODEL_PINHOLE, n, capi.F32)
st.generate(seed=3, gt=M.reshape(-1), lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5)
prob = capi.make_problem(capi.MODEL_PINHOLE, capi.JAC_CENTRAL, capi.F32, consts=consts)
for _ in range(3):
    ctx.linearize(st, prob, [0.0] * 6)
st.close()
x15 = np.concatenate([x_gt, [600.0, 600.0, 320.0, 240.0, 0.05, -0.02, 0.001, -0.001, 0.005]])
st = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n, capi.F32)
st.generate(seed=3, gt=x15, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=consts[12:])
prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F32, consts=consts[12:])
for _ in range(3):
    ctx.linearize(st, prob, x15 * 0.999)
st.close()
n = 10_000_000
st = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
prob = capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F32)
for _ in range(3):
    ctx.linearize(st, prob, [0.25, 0.15])
st.close()
ctx.close()
