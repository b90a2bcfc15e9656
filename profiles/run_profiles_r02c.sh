#!/bin/bash
# Round 2, third call: full GPU suite (fixed tolerances, prediction / training / file-format tests), the PCG history
# diagnostic, and the headline bench with per-step phase times.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02c_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02c_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02c_gpu_tests.log | cut -c1-300 | head -60
timeout 300 python scripts/debug_pcg_hist.py > gpurun_out/r02c_pcg_hist.log 2>&1; stamp "pcg hist rc=$?"
grep -E "it |atol" gpurun_out/r02c_pcg_hist.log; tail -12 gpurun_out/r02c_pcg_hist.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r02c_bench_default.json 2> gpurun_out/r02c_bench_default.err; stamp "bench default rc=$?"
tail -3 gpurun_out/r02c_bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02c_bench_default.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
    for s in d['phases']['per_step']: print(s)
    print({k: d['phases'][k] for k in ('preconditioner_s', 'pchol_build_s', 'assemble_s', 'cg_s', 'cg_iters', 'precon_apply_avg_ms')})
    print('alt', d.get('alt')); print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('detail'))
except Exception as e:
    print('parse failed', e)
PY
