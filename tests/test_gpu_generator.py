"""The device-side synthetic generator of the bench workload (csrc/mopt_store.cu generate_p2p_kernel) and its host
restatement (oracle/oracle_capi.cpp orc_generate_p2p) produce the same streams bit for bit, so `bench.py --impl
reference` times the CPU path on exactly the 100 M correspondences the GPU arm linearizes."""
import numpy as np
import pytest

from oracle import oracle_py as orc

pytestmark = pytest.mark.gpu
X_GT = [0.5, -0.3, 0.2, 0.10, -0.05, 0.08]


@pytest.mark.parametrize("first,n", [(0, 1_000_003), (99_000_000, 500_001), (3_999_999_000, 70_001)])
def test_host_generator_is_bit_identical_to_the_device_generator(first, n):
    from moptimizer_0_b200 import capi
    ctx = capi.Context(0)
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
    kw = dict(lo=(0, 0, 0), hi=(10, 10, 10), first_index=first, noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
    st.generate(seed=2, gt=X_GT, **kw)
    ds, dt = st.download(0, np.float32), st.download(1, np.float32)
    hs, ht = orc.generate_p2p(2, n, X_GT, **kw)
    assert np.array_equal(ds, hs.astype(np.float32)) and np.array_equal(ds.astype(np.float64), hs)
    assert np.array_equal(dt, ht.astype(np.float32)) and np.array_equal(dt.astype(np.float64), ht)
    # ~5 % of the rows carry an outlier offset (both sides agree on which)
    out = np.max(np.abs(ht - (hs @ orc.so3_convert6dof(X_GT)[:3, :3].T + np.array(X_GT[:3]))), axis=1) > 0.1
    assert 0.03 < out.mean() < 0.07
    st.close()
    ctx.close()
