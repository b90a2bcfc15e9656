#!/bin/bash
# Round 2, sixth call (2 GPUs): peer-memory collectives -- sharded-vs-single parity, then cfg2 with and without them.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02f_mg_check_n$N.log 2>&1; stamp "multi_gpu_check (peer) rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|pivots|Woodbury|projected|Error|rror:|assert" gpurun_out/r02f_mg_check_n$N.log | head -30
MLFFPC_PEER=0 NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02f_mg_check_nccl_n$N.log 2>&1; stamp "multi_gpu_check (nccl) rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|Error|rror:|assert" gpurun_out/r02f_mg_check_nccl_n$N.log | head -20
for PEER in 1 0; do
  MLFFPC_PEER=$PEER timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt \
      > gpurun_out/r02f_bench_cfg2_n${N}_peer$PEER.json 2> gpurun_out/r02f_bench_cfg2_n${N}_peer$PEER.err; stamp "bench cfg2 n=$N peer=$PEER rc=$?"
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02f_bench_cfg2_n${N}_peer$PEER.json'))
    print('peer=$PEER', d['collectives'][:40], 'value', d['value'], 'op ms', d['roofline']['avg_launch_ms'], 'apply ms', d['phases']['precon_apply_avg_ms'])
    for s in d['phases']['per_step']: print('  ', s)
except Exception as e:
    print('parse failed', e)
PY
  grep -vE "^\*|OMP_NUM|^$" gpurun_out/r02f_bench_cfg2_n${N}_peer$PEER.err | tail -4
done
MLFFPC_PEER=1 timeout 600 $TR bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --mode matrix_free \
      > gpurun_out/r02f_bench_cfg2_mf_n${N}.json 2> gpurun_out/r02f_bench_cfg2_mf_n${N}.err; stamp "bench cfg2 mf n=$N rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02f_bench_cfg2_mf_n${N}.json')); print('mf value', d['value'], d['phases']['per_step'])"
