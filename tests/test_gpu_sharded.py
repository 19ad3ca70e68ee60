"""Sharded path on real GPUs: 2 ranks (torchrun) over NCCL must reproduce the single-GPU result.
Skipped unless the box has >= 2 GPUs (`gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _bench(args, nproc):
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                "--master-port", "29611"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(nproc)] + args
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_two_rank_linearization_and_lm_match_single_gpu():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    common = ["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--lm", "--strong-total", "8000000"]
    one = _bench(common, 1)
    two = _bench(common, 2)
    # same 8 M global correspondences, sharded 2 x 4 M: sums agree to fp64 rounding of a different order
    for k in ("sum_rtr", "H00", "b0"):
        assert abs(one["check"][k] - two["check"][k]) <= 1e-9 * abs(one["check"][k]), (k, one["check"], two["check"])
    assert one["lm"]["status"] == two["lm"]["status"]
    assert one["lm"]["executed_iterations"] == two["lm"]["executed_iterations"]
    assert one["lm"]["sequence"] == two["lm"]["sequence"]
    assert abs(one["lm"]["x_err_inf"] - two["lm"]["x_err_inf"]) < 1e-9
