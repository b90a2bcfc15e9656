#!/bin/bash
# A/B runs of library options on the cfg2 system (single GPU, matrix-free operator so each run takes seconds).
# usage: run_ab_numerics.sh TOL "opts A" "opts B" ...
set -u
mkdir -p gpurun_out
TOL=$1; shift
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --tol $TOL --mode matrix_free"
i=0
for opt in "$@"; do
  timeout 600 $B $opt > gpurun_out/r01e_ab_$i.json 2> gpurun_out/r01e_ab_$i.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r01e_ab_$i.json'))
    p = d['phases']
    print('[$opt] iters %d cg %.2f s precon %.3f s apply_ms %.3f hist %s' % (p['cg_iters'], p['cg_s'], p['preconditioner_s'], p['precon_apply_avg_ms'], p['rel_resid_every_100_iters']))
except Exception as e:
    print('[$opt] parse failed', e)
PY
  tail -2 gpurun_out/r01e_ab_$i.err
  i=$((i+1))
done
