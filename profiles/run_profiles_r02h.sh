#!/bin/bash
# Round 2, 8 GPUs: sharded-vs-single parity, cfg2 with peer-memory collectives vs NCCL, and the driver-shaped bench line
# with the north-star block (cfg4 n = 270 000 assembled, cfg5 n = 1.26 M matrix-free).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02h_mg_check_n$N.log 2>&1; stamp "multi_gpu_check (peer) rc=$?"
grep -E "MULTI_GPU_CHECK OK rank 0|iters sharded|pivots|Woodbury|projected|Error|rror:|assert" gpurun_out/r02h_mg_check_n$N.log | head -24
for PEER in 1 0; do
  MLFFPC_PEER=$PEER timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --north-star off \
      > gpurun_out/r02h_bench_cfg2_n${N}_peer$PEER.json 2> gpurun_out/r02h_bench_cfg2_n${N}_peer$PEER.err; stamp "bench cfg2 n=$N peer=$PEER rc=$?"
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02h_bench_cfg2_n${N}_peer$PEER.json'))
    print('peer=$PEER', d['collectives'][:30], 'value', d['value'], 'op ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'apply ms', d['phases']['precon_apply_avg_ms'])
    for s in d['phases']['per_step']: print('  ', s)
except Exception as e:
    print('parse failed', e)
PY
  grep -vE "^\*|OMP_NUM|^$" gpurun_out/r02h_bench_cfg2_n${N}_peer$PEER.err | tail -4
done
timeout 1500 $TR bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline \
      > gpurun_out/r02h_bench_default_n${N}.json 2> gpurun_out/r02h_bench_default_n${N}.err; stamp "driver-shaped bench n=$N rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02h_bench_default_n${N}.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'alt', (d.get('alt') or {}).get('value'))
    ns = d.get('north_star', {})
    for k, v in ns.items():
        print(k, json.dumps({kk: v[kk] for kk in v if kk not in ('note',)})[:1500])
except Exception as e:
    print('parse failed', e)
PY
grep -vE "^\*|OMP_NUM|^$" gpurun_out/r02h_bench_default_n${N}.err | tail -6
