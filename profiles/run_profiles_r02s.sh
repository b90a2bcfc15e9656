#!/bin/bash
# Round 2 (4 GPUs): the world size not yet measured this round -- parity check and the driver-shaped cfg2 line.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 300 $TR tests/multi_gpu_check.py > gpurun_out/r02s_mg_check_n$N.log 2>&1; stamp "multi_gpu_check rc=$?"
grep -E "MULTI_GPU_CHECK OK rank 0|iters sharded|Error|rror:|assert" gpurun_out/r02s_mg_check_n$N.log | head -12
timeout 400 $TR bench.py --gpus $N --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02s_bench_default_n$N.json 2> gpurun_out/r02s_bench_default_n$N.err; stamp "bench default n=$N rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02s_bench_default_n$N.json')); print('n=$N value', d['value'], 'e2e', d['e2e']['value'], 'apply ms', d['phases'].get('precon_apply_avg_ms'), 'roofline', d['roofline'].get('frac'), 'alt', (d.get('alt') or {}).get('value'), d['phases']['per_step'])"
