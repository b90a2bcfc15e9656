#!/bin/bash
# Round 2, final build (1 GPU): ncu launch list of the headline command inside the profiler windows (assembly, 4 pivot
# steps, the whole factor phase, 2 CG iterations).  The ncu pass runs only after the same command exited 0 without ncu.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt"
NCU="ncu --clock-control none --profile-from-start off"
WIN="assemble:0:1,pchol:3000:4,woodbury:0:1,pcg:5:2"
timeout 200 $B > gpurun_out/r02v_head_plain.json 2> gpurun_out/r02v_head_plain.err && {
  MLFFPC_PROFILE=$WIN timeout 200 $NCU --metrics gpu__time_duration.sum --csv \
      --log-file gpurun_out/r02v_launches_windows.csv $B --tol 1e-2 > gpurun_out/r02v_launches_stdout.log 2>&1
  echo "launch list rc=$?"
}
python -c "
import json; d=json.load(open('gpurun_out/r02v_head_plain.json')); print('plain value', d['value'], d['phases']['per_step'])"
wc -l gpurun_out/r02v_launches_windows.csv
