"""The persistent one-launch LM kernel (small single-cost point2point problems) against the launch-per-trial loop on 136
small problems (scripts/lm_mono_ab.py): sizes around the CTA / vector granularity, three losses, fp32 / fp64, analytical
and forward-difference Jacobians, both pass orders.  The two paths sum the residuals over different grids, so they agree
to rounding: with fp64 data the solutions coincide; with fp32 data the forward-difference cases wander at their noise
floor (sqrt(eps_f32) relative Jacobian error) and end within it of each other."""
import json, os, subprocess, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "lm_mono_ab.py")], env=env, capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_persistent_lm_kernel_agrees_with_launch_per_trial_loop():
    from moptimizer_0_b200 import capi
    mono, loop = _run({}), _run({"MOPT_LM_MONO": "0"})
    assert len(mono) == len(loop) == 136
    for a, b in zip(mono, loop):
        assert a["case"] == b["case"]
        n, dtype, jac, loss, spec = a["case"]
        dx = float(np.max(np.abs(np.array(a["x"]) - np.array(b["x"]))))
        assert a["status"] in ("CONVERGED", "SMALL_DELTA", "MAXIMUM_ITERATIONS_REACHED"), a
        if dtype == capi.F64:
            assert dx < (1e-9 if jac == capi.JAC_ANALYTICAL else 1e-7) and a["status"] == b["status"], (a, b)
        elif jac == capi.JAC_ANALYTICAL:
            assert dx < 1e-6, (a, b)
        else:
            assert dx < 1e-4, (a, b)
