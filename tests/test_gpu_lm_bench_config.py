"""LM at BASELINE configs[2] (point2point, analytical Jacobian + Huber 0.05, x0 = 0) against the oracle's LM on a
12 M-row slice of the bench generator's data: status, accept/reject sequence, y0 / lambda traces, iteration count and
final parameters (north star: x within 1e-6, same iteration count and sequence).

What "same sequence" can mean (DESIGN.md "LM tail"): rho = (y0 - yi) / den compares two sums of r^T r.  Once the step
is so small that |y0 - yi| / y0 drops to the rounding noise of those sums, the sign of rho is noise in the reference
itself (its y0 and yi are summed in different orders: serial loop, linearization.h:142-154, against TBB
parallel_reduce, :52-62).  The traces are therefore compared trial by trial down to a stated floor of the relative
decrease — 1e-13 for fp64 residual arithmetic, 2e-9 for fp32 — and below it only the outcome is compared: same
status, iteration counts that differ by no more than the number of below-floor trials, same x."""
import numpy as np
import pytest

from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu
X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]
N = 12_000_000
FLOOR_F64, FLOOR_F32 = 1e-13, 2e-9


@pytest.fixture(scope="module")
def setup():
    from moptimizer_0_b200 import capi
    ctx = capi.Context(0)
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, N, capi.F32)
    st.generate(seed=2, gt=X_GT, lo=(0, 0, 0), hi=(10, 10, 10), noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    src, tgt = st.download(0, np.float64), st.download(1, np.float64)
    th = max(2, orc.hardware_concurrency() or 8)
    oc = orc.Cost(orc.P2P, 6, 3, N, a=src, b=tgt, jac_mode=orc.JAC_ANALYTICAL, variant=orc.P2P_EXACT,
                  loss=orc.LOSS_HUBER, loss_param=0.05, lin_threads=th, cost_threads=th)
    ref = orc.lm_minimize([oc], [0.0] * 6, max_iterations=50, stagnation_stop=True)
    ref_literal = orc.lm_minimize([oc], [0.0] * 6, max_iterations=50, stagnation_stop=False)
    yield capi, ctx, st, ref, ref_literal
    st.close()
    ctx.close()


def above_floor(trace, floor):
    """Number of leading trials whose relative decrease is above `floor`."""
    k = 0
    for t in trace:
        if abs(t[2] - t[3]) <= floor * abs(t[2]):
            break
        k += 1
    return k


def compare(dev, ref, floor, x_tol, tag):
    k = min(above_floor(ref.trace, floor), above_floor(dev.trace, floor))
    assert k >= 5, (tag, k, ref.sequence, dev.sequence)           # the descent itself is well above the floor
    assert dev.sequence[:k] == ref.sequence[:k], (tag, dev.sequence, ref.sequence)
    for td, tr in zip(dev.trace[:k], ref.trace[:k]):
        assert (td[0], td[1]) == (tr[0], tr[1])                    # same outer iteration / inner try
        assert abs(td[2] - tr[2]) <= 1e-6 * tr[2] and abs(td[5] - tr[5]) <= 1e-6 * tr[5], (tag, td, tr)  # y0, lambda
    assert dev.status == ref.status == "SMALL_DELTA", (tag, dev.status, ref.status)
    tail = max(len(ref.trace), len(dev.trace)) - k
    assert abs(dev.executed_iterations - ref.executed_iterations) <= tail, (tag, dev.executed_iterations,
                                                                             ref.executed_iterations, tail)
    err = float(np.max(np.abs(dev.x - ref.x)))
    assert err <= x_tol, (tag, err)
    return k, err


def test_lm_fp64_compute_matches_oracle(setup):
    capi, ctx, st, ref, ref_literal = setup
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F64, loss=capi.LOSS_HUBER,
                             loss_param=0.05, variant=capi.P2P_EXACT)
    dev = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50)
    k, err = compare(dev, ref, FLOOR_F64, 1e-9, "f64")
    print(f"f64: {k} trials above the floor identical; oracle {ref.sequence} ({ref.executed_iterations}), "
          f"device {dev.sequence} ({dev.executed_iterations}), |dx| = {err:.1e}")
    # the literal loop (no stagnation stop) agrees over the same prefix on both sides; how it ends is rounding noise
    lit = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50, stagnation_stop=False)
    kk = min(above_floor(ref_literal.trace, FLOOR_F64), above_floor(lit.trace, FLOOR_F64))
    assert kk == k and lit.sequence[:kk] == ref_literal.sequence[:kk]
    assert float(np.max(np.abs(lit.x - ref_literal.x))) <= 1e-9
    # ... and the stop changes the answer by less than isDeltaSmall's threshold
    assert float(np.max(np.abs(lit.x - dev.x))) < 1.5e-8


def test_lm_fp32_compute_matches_oracle_to_its_floor(setup):
    capi, ctx, st, ref, _ = setup
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32, loss=capi.LOSS_HUBER,
                             loss_param=0.05, variant=capi.P2P_EXACT)
    dev = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50)
    k, err = compare(dev, ref, FLOOR_F32, 1e-6, "f32")
    print(f"f32: {k} trials above the floor identical; oracle {ref.sequence} ({ref.executed_iterations}), "
          f"device {dev.sequence} ({dev.executed_iterations}), |dx| = {err:.1e}")
