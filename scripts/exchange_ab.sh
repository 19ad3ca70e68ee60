#!/bin/bash
# A/B of the cross-GPU exchange at N ranks: fused consumer (default) vs separate consumer kernel vs NCCL,
# on a tiny pass (exchange latency) and on the 100 M pass.   usage: scripts/exchange_ab.sh N
N=$1
run() {  # $1 = label, $2 = env assignment, rest = bench args
  label=$1; envs=$2; shift 2
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
    bench.py --gpus $N --no-e2e --no-cpu-baseline --no-lm "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$label', 'n/gpu', d['config']['n_per_gpu'], 'coupled us/step %.1f' % (d['ms_per_step']*1e3), 'uncoupled max %.1f' % (max(d['per_rank_uncoupled_ms_per_step'])*1e3), 'Gres/s %.1f' % d['value'])"
}
for rep in 1 2; do
  run "fused   " "MOPT_X=1" --per-gpu 4096 --steps 500 --prewarm-steps 200
  run "kernel  " "MOPT_PEER_CONSUMER=kernel" --per-gpu 4096 --steps 500 --prewarm-steps 200
  run "nccl    " "MOPT_X=1" --per-gpu 4096 --steps 500 --prewarm-steps 200 --collective nccl
done
run "fused   " "MOPT_X=1" --steps 200
run "kernel  " "MOPT_PEER_CONSUMER=kernel" --steps 200
run "nccl    " "MOPT_X=1" --steps 200 --collective nccl
run "fused   " "MOPT_X=1" --steps 200
