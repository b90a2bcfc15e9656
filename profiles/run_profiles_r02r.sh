#!/bin/bash
# Round 2 (1 GPU): headline bench, 4 timed steps, clock sampling at 1 Hz (is the pivot-loop jitter of the later steps gone?)
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps ${STEPS:-4} --warmup 1 --no-cpu-baseline --no-alt > gpurun_out/${TAG:-r02r}_bench_default.json 2> gpurun_out/${TAG:-r02r}_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/%s_bench_default.json' % __import__("os").environ.get("TAG", "r02r")))
print('value', d['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
for s in d['phases']['per_step']: print('  ', s)
PY
