"""Worker for tests/test_gpu_sharded.py: runs under torchrun (or alone) and prints one JSON line on rank 0.
Every rank holds a contiguous shard; results must not depend on the number of ranks."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from moptimizer_0_b200 import capi, sharding
    from tests.common import fachada

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    want = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    used = "none"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ctx, used, note = sharding.make_sharded_context(local, rank, world, want)
        assert used == want, f"collective {want} unavailable: {note}"
    else:
        ctx = capi.Context(local)
    out = {"collective": used}

    # (1) fachada, fp64 store + fp64 compute, numerical Jacobian: exact LM parity across shardings
    src, tgt, _, _ = fachada()
    lo, hi = sharding.shard_range(src.shape[0], rank, world)
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, hi - lo, capi.F64)
    st.upload(0, src[lo:hi])
    st.upload(1, tgt[lo:hi])
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_FORWARD, capi.F64)
    x = [1.0, 2.0, 3.0, 0.2, -0.3, 0.4]
    H, b, s = ctx.linearize(st, prob, x)
    cost = ctx.compute_cost(st, prob, x)
    r = ctx.lm_minimize([st], [prob], [0.0] * 6, max_iterations=50)
    out["fachada"] = {"H": H.tolist(), "b": b.tolist(), "sum": s, "cost": cost, "status": r.status,
                      "executed": r.executed_iterations, "sequence": r.sequence, "x": r.x.tolist(),
                      "y0": r.trace[:, 2].tolist()}
    st.close()

    # (2) synthetic fp32 workload of the benchmark, Huber loss: global sums independent of the sharding
    n_total = 8_000_003
    lo, hi = sharding.shard_range(n_total, rank, world)
    st = capi.Store(ctx, capi.MODEL_POINT2POINT, hi - lo, capi.F32)
    st.generate(seed=2, gt=[0.5, -0.3, 0.2, 0.10, -0.05, 0.08], first_index=lo, noise_sigma=0.01,
                outlier_fraction=0.05, outlier_range=1.0)
    prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32, loss=capi.LOSS_HUBER,
                             loss_param=0.05)
    H, b, s = ctx.linearize(st, prob, [0.0] * 6)
    out["synthetic"] = {"H": H.tolist(), "b": b.tolist(), "sum": s}
    st.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
