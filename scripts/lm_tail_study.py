#!/usr/bin/env python
"""Where does the accept/reject tail of the point2point LM come from?  

Runs the device LM on the bench workload (Huber 0.05, x0 = 0) at 10 M, 100 M and 800 M correspondences (800 M is
the 8-GPU job's total, here on one GPU) with fp32 and fp64 residual arithmetic, prints every trial
(relative decrease, rho, lambda) and, at 10 M, the oracle's fp64 trace on the same rows.

    python scripts/lm_tail_study.py [sizes in millions ...]          (GPU box)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moptimizer_0_b200 import capi  # noqa: E402

X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]


def show(tag, r, dt):
    print(f"{tag}: {r.status} after {r.executed_iterations} iterations, {len(r.trace)} trials, {dt * 1e3:.1f} ms, "
          f"sequence {r.sequence}")
    print("   x =", np.array2string(np.asarray(r.x), precision=10))
    for t in r.trace:
        y0, yi, rho, lam = t[2], t[3], t[4], t[5]
        print(f"   it {int(t[0]):2d}.{int(t[1])}  y0 {y0:.12e}  (y0-yi)/y0 {(y0 - yi) / y0:+.3e}  rho {rho:+.3e}  "
              f"lambda {lam:.3e}  {'A' if t[7] else 'R'}")


def main():
    sizes = [int(float(a) * 1e6) for a in sys.argv[1:]] or [10_000_000, 100_000_000, 800_000_000]
    ctx = capi.Context(0)
    for n in sizes:
        st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
        st.generate(seed=2, gt=X_GT, lo=(0, 0, 0), hi=(10, 10, 10), noise_sigma=0.01, outlier_fraction=0.05,
                    outlier_range=1.0)
        res = {}
        for name, cd, sd in (("f32 compute / f64 LM", capi.F32, capi.F64), ("f32 compute / f32 LM", capi.F32, capi.F32),
                             ("f64 compute / f64 LM", capi.F64, capi.F64)):
            prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, cd, loss=capi.LOSS_HUBER,
                                     loss_param=0.05, variant=capi.P2P_EXACT)
            ctx.lm_minimize([st], [prob], np.zeros(6), max_iterations=2, scalar_dtype=sd)
            t0 = time.perf_counter()
            r = ctx.lm_minimize([st], [prob], np.zeros(6), max_iterations=50, scalar_dtype=sd)
            show(f"n = {n / 1e6:.0f} M, {name}", r, time.perf_counter() - t0)
            res[name] = r
        a, b = res["f32 compute / f64 LM"], res["f64 compute / f64 LM"]
        print(f"   |x_f32 - x_f64|_inf = {np.max(np.abs(a.x - b.x)):.3e}")
        if n <= 20_000_000:
            from oracle import oracle_py as orc
            orc.build()
            th = orc.hardware_concurrency() or 8
            src, tgt = st.download(0, np.float64), st.download(1, np.float64)
            oc = orc.Cost(orc.P2P, 6, 3, n, a=src, b=tgt, jac_mode=orc.JAC_ANALYTICAL, variant=orc.P2P_EXACT,
                          loss=orc.LOSS_HUBER, loss_param=0.05, lin_threads=th, cost_threads=th)
            t0 = time.perf_counter()
            ro = orc.lm_minimize([oc], [0.0] * 6, max_iterations=50)
            show(f"n = {n / 1e6:.0f} M, ORACLE f64 ({th} threads)", ro, time.perf_counter() - t0)
            print(f"   |x_dev64 - x_oracle|_inf = {np.max(np.abs(b.x - ro.x)):.3e}   |x_dev32 - x_oracle|_inf = "
                  f"{np.max(np.abs(a.x - ro.x)):.3e}")
        st.close()
    ctx.close()


if __name__ == "__main__":
    main()
