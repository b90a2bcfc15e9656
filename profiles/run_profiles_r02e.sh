#!/bin/bash
# Round 2, fifth call (2 GPUs): TMA row-strip GEMV tests, sharded-vs-single parity with the new collectives
# (merged pivot message, fused scalar allreduce, (hi, lo) Gram fold, unsynchronised seeds), cfg2 bench on 2 GPUs.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_precon_forms.py tests/test_gpu_cfg1.py -m gpu -q > gpurun_out/r02e_gpu_tests.log 2>&1; stamp "pytest forms+cfg1 rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02e_gpu_tests.log | tail; grep -E "^E  " gpurun_out/r02e_gpu_tests.log | cut -c1-250 | head -20
for TR in 1 0; do
  timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-alt --mode matrix_free --opt tma_rows=$TR > gpurun_out/r02e_mf_tma$TR.json 2> gpurun_out/r02e_mf_tma$TR.err
  python -c "
import json; d=json.load(open('gpurun_out/r02e_mf_tma$TR.json')); p=d['phases']; print('tma_rows=$TR value %.3f iters %d apply_ms %.4f op_ms %.4f' % (d['value'], p['cg_iters'], p['precon_apply_avg_ms'], d['roofline']['avg_launch_ms']))"
done
stamp "tma rows A/B"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 900 $TR tests/multi_gpu_check.py > gpurun_out/r02e_mg_check_n$N.log 2>&1; stamp "multi_gpu_check rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|pivots|Woodbury|projected|Error|rror:|assert" gpurun_out/r02e_mg_check_n$N.log | head -30
timeout 900 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r02e_bench_cfg2_n$N.json 2> gpurun_out/r02e_bench_cfg2_n$N.err; stamp "bench cfg2 n=$N rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02e_bench_cfg2_n$N.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline frac', d['roofline']['frac'], 'op ms', d['roofline']['avg_launch_ms'])
    for s in d['phases']['per_step']: print('  ', s)
    print('  apply ms', d['phases']['precon_apply_avg_ms'], 'alt', d.get('alt'))
except Exception as e:
    print('parse failed', e)
PY
grep -vE "^\*|OMP_NUM|^$" gpurun_out/r02e_bench_cfg2_n$N.err | tail -5
