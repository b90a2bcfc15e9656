#!/bin/bash
# Round-1 final state: GPU test suite + launch list of the headline command (profiler windows) with the final kernels.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r01r_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
tail -3 gpurun_out/r01r_gpu_tests.log
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt"
WIN="assemble:0:1,pchol:3000:2,syrk:0:1,potrf:10:1,trsm:5:1,pcg:5:2"
timeout 300 $B > gpurun_out/r01r_head_plain.log 2>&1 && {
  stamp "head plain ok"
  MLFFPC_PROFILE=$WIN timeout 600 ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv \
      --log-file gpurun_out/r01r_launches_windows.csv $B > gpurun_out/r01r_launches_stdout.log 2>&1
  stamp "launch list rc=$?"
}
