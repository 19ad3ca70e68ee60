"""Worker of tests/test_gpu_setup_forms.py (run with MOPT_LM_MONO=0): the first pass of a device LM solve takes its
ParamBlock from the warp-parallel set-up of the optimizer step (so3_exp_jl_warp); a host-driven linearize at the same x
takes it from the serial forms (fused into the pass kernel, or the set-up kernel with MOPT_FUSED_SETUP=0).  Prints, per
x, the bits of (sum, max diag H) from both."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi

ctx = capi.Context(0)
n = 20_000
out = []
for dtype in (capi.F64, capi.F32):
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, dtype)
    st.generate(seed=11, gt=[0.5, -0.3, 0.2, 0.10, -0.05, 0.08], lo=(2, -14, -2), hi=(20, 8, 6), noise_sigma=0.01)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, dtype, variant=capi.P2P_EXACT)
    for x0 in ([0.0] * 6, [0.1, 0.2, 0.3, 1e-5, -2e-5, 3e-5], [0.4, -0.2, 0.1, 0.1, -0.05, 0.08],
               [1.0, 2.0, 3.0, 2.5, -1.0, 0.7], [0.0, 0.0, 0.0, 1e-17, 0.0, 0.0]):
        H, b, s = ctx.linearize(st, prob, x0)
        r = ctx.lm_minimize([st], [prob], x0, max_iterations=1, lm_iterations=1)
        out.append({"dtype": dtype, "x0": x0, "sum_linearize": float(s).hex(), "sum_lm": float(r.trace[0, 2]).hex(),
                    "lambda_linearize": float(1e-9 * np.max(np.abs(np.diag(H)))).hex(),
                    "lambda_lm": float(r.trace[0, 5]).hex()})
    st.close()
# exp-curve, finite differences, fp32: the sets are built by the set-up kernel (MOPT_FUSED_SETUP=0) or by every CTA of
# the pass kernel itself (default) — the caller compares the two runs' bits
curve = []
st = capi.Store(ctx, capi.MODEL_EXP_CURVE, 100_003, capi.F32)
st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=100_003, noise_sigma=0.2)
for jac in (capi.JAC_FORWARD, capi.JAC_CENTRAL):
    prob = capi.make_problem(capi.MODEL_EXP_CURVE, jac, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.3)
    for x0 in ([0.0, 0.0], [0.25, 0.15], [-1.5, 2.0]):
        H, b, s = ctx.linearize(st, prob, x0)
        c = ctx.compute_cost(st, prob, x0)
        curve.append([float(v).hex() for v in list(H.reshape(-1)) + list(b) + [s, c]])
st.close()
print(json.dumps({"p2p": out, "curve": curve}))
