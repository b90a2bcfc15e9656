#!/bin/bash
# Round 2 (N GPUs): peer flags with one fence per signal / wait (relaxed stores and polls): parity, cfg2 bench with the
# one-launch tile pass on / off, pivot-loop time split.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 400 $TR tests/multi_gpu_check.py > gpurun_out/r02o_mg_check_n$N.log 2>&1; stamp "multi_gpu_check rc=$?"
grep -E "MULTI_GPU_CHECK OK rank 0|iters sharded|Error|rror:|assert" gpurun_out/r02o_mg_check_n$N.log | head -12
for SM in 1 0; do
  MLFFPC_TIMING=1 timeout 400 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --north-star off --opt symop_multi=$SM > gpurun_out/r02o_bench_cfg2_n${N}_multi$SM.json 2> gpurun_out/r02o_bench_cfg2_n${N}_multi$SM.err; stamp "bench cfg2 n=$N symop_multi=$SM rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_cfg2_n${N}_multi$SM.json')); print('n=$N multi=$SM value', d['value'], 'apply ms', d['phases'].get('precon_apply_avg_ms'), 'roofline', d['roofline'].get('frac'), d['phases']['per_step'])"
  grep "pchol look-ahead" gpurun_out/r02o_bench_cfg2_n${N}_multi$SM.err | tail -1
done
