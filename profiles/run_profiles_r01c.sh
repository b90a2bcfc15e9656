#!/bin/bash
# Round-1 GPU evidence, third pass: look-ahead pivoted Cholesky, preconditioner-accuracy A/B, symv_tma capture, cfg1.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
NCU="ncu --clock-control none --profile-from-start off"

timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01c_gpu_parity.log 2>&1; stamp "parity rc=$?"
tail -5 gpurun_out/r01c_gpu_parity.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r01c_gpu_fullsize.log 2>&1; stamp "fullsize rc=$?"
tail -5 gpurun_out/r01c_gpu_fullsize.log

i=0
for opt in "" "--opt pchol_lookahead=0" "--opt precon_accuracy=2" "--opt precon_accuracy=1"; do
  timeout 600 $B --tol 1e-3 --mode matrix_free $opt > gpurun_out/r01c_ab_$i.json 2> gpurun_out/r01c_ab_$i.err; stamp "A/B $i ($opt) rc=$?"
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r01c_ab_$i.json'))
    p = d['phases']
    print('   value %.2f pchol %.3f precon %.3f cg %.2f iters %d apply_ms %.3f op_ms %.3f' % (d['value'], p['pchol_build_s'], p['preconditioner_s'], p['cg_s'], p['cg_iters'], p['precon_apply_avg_ms'], d['roofline']['avg_launch_ms']))
except Exception as e:
    print('   parse failed', e)
PY
  tail -2 gpurun_out/r01c_ab_$i.err
  i=$((i+1))
done

timeout 600 $B --workload cfg1 --mode assembled_sym > gpurun_out/r01c_cfg1.json 2> gpurun_out/r01c_cfg1.err; stamp "cfg1 rc=$?"
tail -c 700 gpurun_out/r01c_cfg1.json; tail -3 gpurun_out/r01c_cfg1.err

S="$B --tol 1e-2 --mode assembled_sym"
timeout 300 $S > gpurun_out/r01c_sym_plain.log 2>&1 && {
  MLFFPC_PROFILE=pcg:5:1 timeout 600 $NCU --set full --import-source on -k regex:'symv_tma' -c 2 \
      -o gpurun_out/r01c_symv_tma -f $S > gpurun_out/r01c_symv_stdout.log 2>&1
  stamp "symv_tma capture rc=$?"
}
ls -la gpurun_out/ | tail -8
