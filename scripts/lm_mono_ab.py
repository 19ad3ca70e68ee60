"""Persistent LM kernel against the launch-per-trial loop (MOPT_LM_MONO=0) on a spread of small point2point problems:
sizes around the CTA / vector granularity, the three losses, fp32 and fp64, analytical and finite-difference Jacobians,
speculative and reference pass order.  Prints one JSON line; run it under both settings and compare (scripts/lm_mono_ab.py
--compare a.json b.json)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) == 4 and sys.argv[1] == "--compare":
    a, b = (json.load(open(p)) for p in sys.argv[2:4])
    worst = 0.0
    assert len(a) == len(b)
    for ra, rb in zip(a, b):
        assert ra["case"] == rb["case"]
        same = ra["status"] == rb["status"] and ra["sequence"] == rb["sequence"]
        dx = float(np.max(np.abs(np.array(ra["x"]) - np.array(rb["x"]))))
        worst = max(worst, dx)
        if not same or dx > 1e-6:
            print("DIFF", ra["case"], ra["status"], rb["status"], ra["sequence"], rb["sequence"], dx)
    print(f"{len(a)} cases, worst |x_mono - x_launch_per_trial| = {worst:.3e}")
    sys.exit(0)

from moptimizer_0_b200 import capi
ctx = capi.Context(0)
X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]
out = []
for n in (1, 3, 255, 1024, 5_000, 29_310, 200_001):
    for dtype in (capi.F32, capi.F64):
        st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, dtype)
        st.generate(seed=7 + n, gt=X_GT, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
        for jac in (capi.JAC_ANALYTICAL, capi.JAC_FORWARD):
            for loss, lp in ((capi.LOSS_NONE, 0.0), (capi.LOSS_HUBER, 0.05), (capi.LOSS_GEMAN_MCCLURE, 0.5)):
                for spec in (True, False):
                    if n < 6 and loss != capi.LOSS_NONE:
                        continue
                    prob = capi.make_problem(capi.MODEL_POINT2POINT, jac, dtype, loss=loss, loss_param=lp, variant=capi.P2P_EXACT)
                    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=12, speculative=spec)
                    out.append({"case": [n, dtype, jac, loss, spec], "status": r.status, "sequence": r.sequence,
                                "x": [float(v) for v in r.x], "passes": r.num_passes})
        st.close()
print(json.dumps(out))
