#!/bin/bash
# Round 2, first call: iteration counts of the existing preconditioner forms on cfg1 (reference: 119 CG iterations
# = num_iters 120) and on cfg2 through the matrix-free operator, before any kernel changes.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
run() { # name, args...
  local name=$1; shift
  timeout 900 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt "$@" > gpurun_out/r02a_$name.json 2> gpurun_out/r02a_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02a_$name.json'))
    p = d['phases']
    print('[$name] iters %d value %.3f s cg %.3f s precon %.3f s apply_ms %.3f conv %s rel %.3e' % (p['cg_iters'], d['value'], p['cg_s'], p['preconditioner_s'], p['precon_apply_avg_ms'], p['converged'], p['rel_resid']))
except Exception as e:
    print('[$name] parse failed', e)
PY
  echo "[$(( $(date +%s) - T0 )) s]"
}
run cfg1_woodbury --workload cfg1 --precon-form woodbury
run cfg1_woodbury_chunk256 --workload cfg1 --precon-form woodbury --opt syrk_chunk=256
run cfg1_woodbury_chunk64 --workload cfg1 --precon-form woodbury --opt syrk_chunk=64
run cfg1_orthonormal --workload cfg1 --no-reorth
run cfg1_orthonormal_chunk256 --workload cfg1 --no-reorth --opt syrk_chunk=256
run cfg1_reorth --workload cfg1
run cfg1_woodbury_mf --workload cfg1 --precon-form woodbury --mode matrix_free
run cfg2_woodbury_chunk256_mf --workload cfg2 --precon-form woodbury --opt syrk_chunk=256 --mode matrix_free
run cfg2_woodbury_chunk64_mf --workload cfg2 --precon-form woodbury --opt syrk_chunk=64 --mode matrix_free
run cfg2_reorth_mf --workload cfg2 --mode matrix_free
nvidia-smi --query-gpu=name,memory.total --format=csv
