#!/bin/bash
# Round 2, closing single-GPU evidence: suite, smoke, headline bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG:-r02q}_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/${TAG:-r02q}_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/${TAG:-r02q}_gpu_tests.log | cut -c1-300 | head -40
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG:-r02q}_smoke.log 2>&1; stamp "smoke rc=$?"
tail -5 gpurun_out/${TAG:-r02q}_smoke.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/${TAG:-r02q}_bench_default.json 2> gpurun_out/${TAG:-r02q}_bench_default.err; stamp "bench default rc=$?"
python - <<'PY'
import json
try:
    import os
    d = json.load(open('gpurun_out/%s_bench_default.json' % os.environ.get('TAG', 'r02q')))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'alt', d.get('alt'))
    for s in d['phases']['per_step']: print('  ', s)
    print('roofline', d['roofline'])
except Exception as e:
    print('parse failed', e)
PY
