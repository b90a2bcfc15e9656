#!/bin/bash
# Round 2, second call: GPU test suite with the extended-precision Gram, the projected form and the device-driven PCG
# loop; iteration counts of every form on cfg1 / cfg2; first headline line with the two-pass projected apply.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
tail -40 gpurun_out/r02b_gpu_tests.log
run() { # name, args...
  local name=$1; shift
  timeout 900 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt "$@" > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02b_$name.json'))
    p = d['phases']
    print('[$name] iters %d value %.3f s cg %.3f s precon %.3f s (pchol %.3f) apply_ms %.3f op_ms %.3f conv %s rel %.3e' % (p['cg_iters'], d['value'], p['cg_s'], p['preconditioner_s'], p['pchol_build_s'], p['precon_apply_avg_ms'], d['roofline']['avg_launch_ms'], p['converged'], p['rel_resid']))
except Exception as e:
    print('[$name] parse failed', e)
PY
  tail -2 gpurun_out/r02b_$name.err
  stamp "$name"
}
run cfg1_woodbury --workload cfg1 --precon-form woodbury
run cfg1_projected --workload cfg1 --precon-form projected
run cfg2_projected_mf --workload cfg2 --precon-form projected --mode matrix_free
run cfg2_projected_exact_mf --workload cfg2 --precon-form projected --mode matrix_free --opt defect_mode=2
run cfg2_woodbury_mf --workload cfg2 --precon-form woodbury --mode matrix_free
run cfg2_reorth_mf --workload cfg2 --precon-form reorth --mode matrix_free
run cfg2_projected_mf_2 --workload cfg2 --precon-form projected --mode matrix_free
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err; stamp "bench default rc=$?"
tail -3 gpurun_out/r02b_bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02b_bench_default.json'))
    print(json.dumps({k: d[k] for k in ('value', 'e2e', 'roofline', 'phases', 'alt', 'gpu_launches') if k in d}, indent=1)[:3500])
    print(d.get('cpu_baseline', {}).get('detail'))
except Exception as e:
    print('parse failed', e)
PY
