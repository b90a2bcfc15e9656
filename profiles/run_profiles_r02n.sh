#!/bin/bash
# Round 2 (8 GPUs): sharded-vs-single parity, rank sweep of the matrix-free north-star system (cfg5), and the
# driver-shaped default bench (cfg2 strong scaling + north-star block) with the factor / pivot-loop time split.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02n_mg_check_n$N.log 2>&1; stamp "multi_gpu_check rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|Error|rror:|assert" gpurun_out/r02n_mg_check_n$N.log | head -24
timeout 900 $TR scripts/cfg5_rank_sweep.py --k 6144 8192 12288 > gpurun_out/r02n_cfg5_sweep.jsonl 2> gpurun_out/r02n_cfg5_sweep.err; stamp "cfg5 rank sweep rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02n_cfg5_sweep.jsonl'):
    try:
        d = json.loads(l)
    except Exception:
        continue
    if 'error' in d: print(d); continue
    p = d['phases']
    print('k', d['k'], 'value', round(d['value'], 2), 'iters', d['cg_iters'], 'conv', d['converged'], 'precon', round(p['preconditioner_s'], 2),
          'pchol', round(p['pchol_build_s'], 2), 'cg', round(p['cg_s'], 2), 'op ms', round(p['operator_avg_ms'], 2), 'apply ms', round(p['precon_apply_avg_ms'], 2),
          'TF/s/GPU', round(d['roofline']['achieved_per_gpu'], 1))
PY
grep -vE "^\*|OMP_NUM|^$|timing" gpurun_out/r02n_cfg5_sweep.err | tail -5
MLFFPC_TIMING=1 timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 1 --no-cpu-baseline --ns-cfg5-k 8192 > gpurun_out/r02n_bench_default_n$N.json 2> gpurun_out/r02n_bench_default_n$N.err; stamp "bench default n=$N rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02n_bench_default_n$N.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'collectives', d.get('collectives'))
    for s in d['phases']['per_step']: print('  ', s)
    print('  matvec ms', d['phases'].get('matvec_avg_ms'), 'apply ms', d['phases'].get('precon_apply_avg_ms'), 'roofline', d['roofline'].get('frac'))
    for kk, v in (d.get('north_star') or {}).items():
        if 'error' in v: print(kk, v); continue
        print(kk, 'value', round(v['value'], 2), 'iters', v['cg_iters'], 'conv', v['converged'], {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v['phases'].items() if a != 'rel_resid_every_200_iters'}, 'roofline frac', round(v['roofline']['frac'], 3))
except Exception as e:
    print('parse failed', e)
PY
grep -E "mlffpc timing" gpurun_out/r02n_bench_default_n$N.err | head -12
