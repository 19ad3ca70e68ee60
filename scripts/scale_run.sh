#!/bin/bash
# Multi-GPU scaling runs (weak: 100 M correspondences per GPU; strong: BASELINE configs[3], 1e9 in total).
# usage: scripts/scale_run.sh "8 4 2" ; writes gpurun_out/scale_n<N>.log
mkdir -p gpurun_out
for N in $1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 100 --warmup 5 --lm --e2e-steps 2 > gpurun_out/scale_n$N.log 2>&1
  echo "N=$N rc=$?"; tail -1 gpurun_out/scale_n$N.log | cut -c1-400
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --collective nccl > gpurun_out/scale_n${N}_nccl.log 2>&1
  echo "N=$N nccl rc=$?"; tail -1 gpurun_out/scale_n${N}_nccl.log | cut -c1-300
done
if [[ " $1 " == *" 8 "* ]]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
    bench.py --gpus 8 --steps 100 --warmup 5 --lm --no-e2e --strong-total 1000000000 > gpurun_out/scale_strong1e9_n8.log 2>&1
  echo "strong rc=$?"; tail -1 gpurun_out/scale_strong1e9_n8.log | cut -c1-400
fi
