"""Launch-shape sweep of the fp32 finite-difference dense pass kernel (needs a build with
MOPT_EXTRA_NVCC_FLAGS=-DMOPT_TUNE_DENSE): camera 50 M central/forward, point2point 100 M forward, curve 10 M central."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moptimizer_0_b200 import capi
from oracle import oracle_py as orc
from tests.common import camera_consts

ctx = capi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())


def time_pass(store, prob, x, steps=10, warm=5):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    for _ in range(warm):
        ctx.linearize_async(store, prob, x)
    ctx.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        ctx.linearize_async(store, prob, x)
    b.record(stream)
    ctx.synchronize()
    return a.elapsed_time(b) / steps


SHAPES = [(0, 0), (2, 0), (3, 0), (4, 0), (4, 128), (6, 128)]
consts = camera_consts()
x_gt = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027])
M = consts[:12].reshape(3, 4) @ orc.so3_convert6dof(x_gt) @ consts[12:].reshape(4, 4)
n = 50_000_000
st = capi.Store(ctx, capi.MODEL_PINHOLE, n, capi.F32)
st.generate(seed=3, gt=M.reshape(-1), lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5)
ref = {}
for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_FORWARD, "forward")):
    prob = capi.make_problem(capi.MODEL_PINHOLE, jac, capi.F32, consts=consts)
    for c, t in SHAPES:
        ctx.set_launch(c, t)
        ms = time_pass(st, prob, [0.0] * 6)
        H, b, s = ctx.result(6)
        ref.setdefault(jn, (H, b, s))
        same = bool(np.allclose(H, ref[jn][0], rtol=1e-6) and np.allclose(b, ref[jn][1], rtol=1e-6, atol=1e-3))
        print(json.dumps({"case": f"camera50M_{jn}_f32", "ctas_per_sm": c, "threads": t or 256, "ms": ms,
                          "Gres_per_s": n / ms / 1e6, "consistent": same}), flush=True)
st.close()
n = 100_000_000
st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
st.generate(seed=2, gt=[0.5, -0.3, 0.2, 0.10, -0.05, 0.08], noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_FORWARD, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.05)
for c, t in SHAPES:
    ctx.set_launch(c, t)
    ms = time_pass(st, prob, [0.0] * 6)
    print(json.dumps({"case": "p2p100M_forward_f32", "ctas_per_sm": c, "threads": t or 256, "ms": ms,
                      "Gres_per_s": n / ms / 1e6}), flush=True)
st.close()
n = 10_000_000
st = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
prob = capi.make_problem(capi.MODEL_EXP_CURVE, capi.JAC_CENTRAL, capi.F32)
for c, t in SHAPES:
    ctx.set_launch(c, t)
    ms = time_pass(st, prob, [0.25, 0.15], steps=50, warm=20)
    print(json.dumps({"case": "curve10M_central_f32", "ctas_per_sm": c, "threads": t or 256, "ms": ms,
                      "Gres_per_s": n / ms / 1e6}), flush=True)
st.close()
# the n x n calibration case (P = 15, wide kernel)
n = 50_000_000
C44 = consts[12:]
x15 = np.concatenate([x_gt, [600.0, 600.0, 320.0, 240.0, 0.05, -0.02, 0.001, -0.001, 0.005]])
st = capi.Store(ctx, capi.MODEL_PINHOLE_DISTORT, n, capi.F32)
st.generate(seed=3, gt=x15, lo=(2.0, -1.0, -0.5), hi=(5.0, 1.0, 1.0), noise_sigma=0.5, consts=C44)
ctx.set_launch(0, 0)
for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_FORWARD, "forward")):
    for cd, cn in ((capi.F32, "f32"), (capi.F64, "f64")):
        prob = capi.make_problem(capi.MODEL_PINHOLE_DISTORT, jac, cd, consts=C44)
        ms = time_pass(st, prob, x15 * 0.999, steps=5, warm=2)
        print(json.dumps({"case": f"camera15_50M_{jn}_{cn}", "ms": ms, "Gres_per_s": n / ms / 1e6}), flush=True)
st.close()
ctx.close()
