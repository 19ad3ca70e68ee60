#!/bin/bash
# A/B of the fused setup on a sharded context (N ranks): default (setup kernel + pass kernel) vs MOPT_FUSED_SETUP=1.
# usage: scripts/fused_setup_ab.sh N   -> gpurun_out/fused_ab_N.log
N=${1:-2}
OUT=gpurun_out/fused_ab_$N.log
: > $OUT
run() {
  echo "== MOPT_FUSED_SETUP=$1" >> $OUT
  MOPT_FUSED_SETUP=$1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 \
    bench.py --gpus $N --steps 200 --warmup 5 --prewarm-steps 1000 --no-e2e --no-cpu-baseline --no-lm 2>/dev/null | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d.get('per_rank_uncoupled_ms_per_step'), d['gpu_launches'])" >> $OUT
}
run 1 29601
run "" 29602
run 1 29603
run "" 29604
cat $OUT
