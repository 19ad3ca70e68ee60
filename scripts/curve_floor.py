"""Exp-curve fit, fp32: time per host-driven linearize step at 1 M / 10 M / 40 M samples, central differences and
analytical Jacobian — the intercept is the fixed cost of a step (launch, prologue, grid reduction), the slope the
streaming rate (~9 us + 3.0 us per million samples for central differences)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moptimizer_0_b200 import capi
ctx = capi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
for n in (1_000_000, 10_000_000, 40_000_000):
    st = capi.Store(ctx, capi.MODEL_EXP_CURVE, n, capi.F32)
    st.generate(seed=1, gt=[0.3, 0.1], lo=(0, 0, 0), hi=(5, 0, 0), n_total=n, noise_sigma=0.2)
    x = np.array([0.25, 0.15])
    for jac, jn in ((capi.JAC_CENTRAL, "central"), (capi.JAC_ANALYTICAL, "analytical")):
        prob = capi.make_problem(capi.MODEL_EXP_CURVE, jac, capi.F32)
        for _ in range(20): ctx.linearize_async(st, prob, x)
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(50): ctx.linearize_async(st, prob, x)
        e1.record(stream); ctx.synchronize()
        print(f"n = {n} {jn}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per step", flush=True)
    st.close()
