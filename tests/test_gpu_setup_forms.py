"""model->setup(x) of the point2point model exists in three forms on the device — serial inside the pass kernel
(fused), serial in the set-up kernel, and spread over a warp in the optimizer step (one sqrt / sincos / division
sequence for so3::Exp and the left Jacobian together).  They must give the same bits: src/so3.cpp:43-57 once."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, MOPT_LM_MONO="0", **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "setup_forms_worker.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("fused", ["1", "0"])
def test_warp_setup_matches_serial_forms_bit_for_bit(fused):
    rows = _run({"MOPT_FUSED_SETUP": fused})["p2p"]
    assert len(rows) == 10
    for r in rows:
        # sum depends on (R, t); lambda_0 = 1e-9 max diag H on a rotational entry, i.e. on the left Jacobian
        assert r["sum_lm"] == r["sum_linearize"], r
        assert r["lambda_lm"] == r["lambda_linearize"], r


@pytest.mark.gpu
def test_fused_finite_difference_setup_matches_setup_kernel_bit_for_bit():
    """Exp curve, fp32 finite differences: the 1 + 2P sets built inside the pass kernel (PassArgs::fused_setup) and by
    the set-up kernel in front of it are the same function on the same x."""
    fused, unfused = _run({})["curve"], _run({"MOPT_FUSED_SETUP": "0"})["curve"]
    assert len(fused) == 6 and fused == unfused
