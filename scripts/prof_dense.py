"""A few launches of the pass kernels of every BASELINE configuration for ncu: point2point 100 M (Huber, analytical),
camera 50 M P=6 dense, P=15 wide, curve 10 M (central differences, fp32).  scripts/ncu_inst_counts.py turns the
capture into profiles/kernel_inst_counts.json (warp instructions per residual, DRAM bytes per launch)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi
from oracle import oracle_py as orc
from tests.common import camera_consts

ctx = capi.Context(0)
REPS = 1 if "--once" in sys.argv else 3  # --once: one launch per kernel (keeps an `ncu --set full` report small)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if len(args) >= 2:
    ctx.set_launch(int(args[0]), int(args[1]))
n = 100_000_000
st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
st.generate(seed=2, gt=[0.5, -0.3, 0.2, 0.10, -0.05, 0.08], lo=(0, 0, 0), hi=(10, 10, 10), noise_sigma=0.01,
            outlier_fraction=0.05, outlier_range=1.0)
prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.05,
                         variant=capi.P2P_EXACT)
for _ in range(REPS):
    ctx.linearize(st, prob, [0.0] * 6)
st.close()
consts = camera_consts()
x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027])
M = consts[:12].reshape(3, 4) @ orc.so3_convert6dof(x_gt) @ consts[12:].reshape(4, 4)
n = 50_000_000
st = capi.Store(ctx, capi.MODEL_PINHOLE, n, capi.F32)
st.generate(seed=3, gt=M.reshape(-1), lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5)
prob = capi.make_problem(capi.MODEL_PINHOLE, capi.JAC_CENTRAL, capi.F32, consts=consts)
for _ in range(REPS):
    ctx.linearize(st, prob, [0.0] * 6)
st.close()
x15 = np.concatenate([x_gt, [600.0, 600.0, 320.0, 240.0, 0.05, -0.02, 0.001, -0.001, 0.005]])
st = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n, capi.F32)
st.generate(seed=3, gt=x15, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=consts[12:])
prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL, capi.F32, consts=consts[12:])
for _ in range(REPS):
    ctx.linearize(st, prob, x15 * 0.999)
st.close()
n = 10_000_000
st = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
prob = capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F32)
for _ in range(REPS):
    ctx.linearize(st, prob, [0.25, 0.15])
st.close()
ctx.close()
