"""Host-side helpers for sharded (multi-GPU) problems: contiguous residual ranges per rank and the packed
(H upper, b, sum) vector that one fp64 all-reduce combines.  Pure Python/numpy so the
world_size>1 logic is testable on CPU with the gloo backend."""
from __future__ import annotations

import numpy as np


def make_sharded_context(local_device: int, rank: int, world: int, collective: str = "p2p"):
    """One mopt context per rank of a torch.distributed job (already initialised).
    collective = "p2p": NVLink peer exchange fused into the pass kernels (CUDA IPC handles gathered over
    torch.distributed); "nccl": one ncclAllReduce per pass.  Returns (ctx, collective actually used, note)."""
    import torch.distributed as dist
    from . import capi
    note = ""
    if collective == "p2p":
        # Every rank runs the SAME sequence of collectives whatever fails locally: gather (ok, handle) first, open the
        # peers only if every rank has a handle, then gather the outcome of the open.
        ctx, mine, ok = None, None, True
        try:
            ctx = capi.Context(local_device, sharded=(rank, world, None))
            mine = ctx.peer_handle()
        except Exception as e:  # e.g. CUDA IPC not permitted in this container
            ok, note = False, str(e)
        got = [None] * world
        dist.all_gather_object(got, (ok, mine))
        if all(g[0] for g in got):
            try:
                ctx.open_peers([g[1] for g in got])
            except Exception as e:
                ok, note = False, str(e)
        else:
            ok = False
        opened = [None] * world
        dist.all_gather_object(opened, ok)
        if all(opened):
            return ctx, "p2p", ""
        if ctx is not None:
            ctx.close()
        note = note or "a peer could not open the exchange"
    uid = [capi.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return capi.Context(local_device, sharded=(rank, world, uid[0])), "nccl", note


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, last) of `rank`: [rank*N/W, (rank+1)*N/W) in integer arithmetic."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def packed_size(P: int) -> int:
    return P * (P + 1) // 2 + P + 1


def pack(H: np.ndarray, b: np.ndarray, s: float) -> np.ndarray:
    """Same layout as csrc PassResult: H upper triangle row-major, then b, then sum."""
    P = b.shape[0]
    iu = np.triu_indices(P)
    return np.concatenate([np.asarray(H, dtype=np.float64)[iu], np.asarray(b, dtype=np.float64), [float(s)]])


def unpack(v: np.ndarray, P: int):
    H = np.zeros((P, P), dtype=np.float64)
    iu = np.triu_indices(P)
    nh = P * (P + 1) // 2
    H[iu] = v[:nh]
    H = H + np.triu(H, 1).T
    return H, np.array(v[nh:nh + P]), float(v[nh + P])
