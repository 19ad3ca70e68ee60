#!/bin/bash
# Final-code multi-GPU runs on one 8-GPU box: weak scaling at 8 and 4 ranks, strong 1e9 at 8, consumer A/B, tiny-pass latency.
mkdir -p gpurun_out
tr() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@"; }
tr 8 --steps 200 --warmup 5 --e2e-steps 2 > gpurun_out/final_n8.log 2>&1; echo "n8 rc=$?"
MOPT_PEER_CONSUMER=kernel tr 8 --steps 200 --warmup 5 --no-e2e --no-lm > gpurun_out/final_n8_kernel_consumer.log 2>&1; echo "n8 kernel-consumer rc=$?"
tr 8 --steps 200 --warmup 5 --no-e2e --strong-total 1000000000 > gpurun_out/final_strong1e9_n8.log 2>&1; echo "strong rc=$?"
tr 4 --steps 200 --warmup 5 --no-e2e > gpurun_out/final_n4.log 2>&1; echo "n4 rc=$?"
tr 8 --per-gpu 4096 --steps 500 --prewarm-steps 200 --no-e2e --no-lm > gpurun_out/final_tiny_n8.log 2>&1; echo "tiny rc=$?"
MOPT_PEER_CONSUMER=kernel tr 8 --per-gpu 4096 --steps 500 --prewarm-steps 200 --no-e2e --no-lm > gpurun_out/final_tiny_n8_kernel.log 2>&1; echo "tiny kernel rc=$?"
tr 8 --per-gpu 4096 --steps 500 --prewarm-steps 200 --no-e2e --no-lm --collective nccl > gpurun_out/final_tiny_n8_nccl.log 2>&1; echo "tiny nccl rc=$?"
for f in gpurun_out/final_*.log; do echo "== $f"; tail -1 $f | cut -c1-200; done
