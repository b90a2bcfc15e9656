#!/bin/bash
# Round-1 closing pass (one GPU): full GPU test suite, smoke(), the headline command and the reference arm.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01n_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
tail -3 gpurun_out/r01n_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01n_smoke.log 2>&1; stamp "smoke rc=$?"
tail -4 gpurun_out/r01n_smoke.log
timeout 1500 python bench.py > gpurun_out/r01n_bench_default.json 2> gpurun_out/r01n_bench_default.err; stamp "default bench rc=$?"
tail -3 gpurun_out/r01n_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/r01n_bench_default.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'phases', d['phases']['preconditioner_s'], d['phases']['assemble_s'], d['phases']['cg_s'], d['phases']['cg_iters'], 'alt', d.get('alt'))"
