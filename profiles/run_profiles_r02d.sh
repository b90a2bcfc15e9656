#!/bin/bash
# Round 2, fourth call (1 GPU): suite with the host-free pivot loop; factor-phase timing breakdown; headline bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02d_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02d_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02d_gpu_tests.log | cut -c1-300 | head -40
timeout 300 python scripts/debug_pcg_hist.py > gpurun_out/r02d_pcg_hist.log 2>&1; stamp "pcg hist rc=$?"
grep -E "it |atol" gpurun_out/r02d_pcg_hist.log; tail -8 gpurun_out/r02d_pcg_hist.log
MLFFPC_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-alt > gpurun_out/r02d_bench_timing.json 2> gpurun_out/r02d_bench_timing.err; stamp "bench timing rc=$?"
grep "mlffpc timing" gpurun_out/r02d_bench_timing.err | tail -24
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err; stamp "bench default rc=$?"
tail -3 gpurun_out/r02d_bench_default.err
python - <<'PY'
import json
for f in ('gpurun_out/r02d_bench_timing.json', 'gpurun_out/r02d_bench_default.json'):
    try:
        d = json.load(open(f))
        print(f, 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
        for s in d['phases']['per_step']: print('  ', s)
        print('  alt', d.get('alt')); print('  cpu', d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print('parse failed', e)
PY
timeout 300 python bench.py --steps 1 --warmup 1 --workload cfg1 --no-e2e --no-cpu-baseline --no-alt > gpurun_out/r02d_bench_cfg1.json 2> gpurun_out/r02d_bench_cfg1.err; stamp "bench cfg1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_cfg1.json')); print('cfg1 value', d['value'], d['phases']['per_step'])"
