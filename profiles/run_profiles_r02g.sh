#!/bin/bash
# Round 2 ncu evidence (1 GPU): launch list of the headline command inside the profiler windows, and --set full captures
# of the kernels of one CG iteration (symmetric operator, preconditioner apply) and of the factor phase (extended-
# precision Gram).  Every ncu pass runs only after the same command has exited 0 without ncu.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt"
NCU="ncu --clock-control none --profile-from-start off"
WIN="assemble:0:1,pchol:3000:4,woodbury:0:1,pcg:5:2"
timeout 300 $B > gpurun_out/r02g_head_plain.json 2> gpurun_out/r02g_head_plain.err && {
  stamp "head plain ok"
  MLFFPC_PROFILE=$WIN timeout 900 $NCU --metrics gpu__time_duration.sum --csv \
      --log-file gpurun_out/r02g_launches_windows.csv $B > gpurun_out/r02g_launches_stdout.log 2>&1
  stamp "launch list rc=$?"
  MLFFPC_PROFILE=pcg:5:1 timeout 900 $NCU --set full --import-source on -c 12 \
      -o gpurun_out/r02g_pcg_iteration -f $B --tol 1e-2 > gpurun_out/r02g_pcg_stdout.log 2>&1
  stamp "pcg iteration capture rc=$?"
  MLFFPC_PROFILE=woodbury:0:1 timeout 900 $NCU --set full --import-source on -k regex:'gram_dd' -c 3 \
      -o gpurun_out/r02g_gram_dd -f $B --tol 1e-2 > gpurun_out/r02g_gram_stdout.log 2>&1
  stamp "gram_dd capture rc=$?"
}
ls -la gpurun_out/ | grep r02g
