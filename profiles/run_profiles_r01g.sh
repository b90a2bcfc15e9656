#!/bin/bash
# Round-1 final evidence pass (one GPU): the headline command as the driver runs it, its launch list inside the
# profiler windows, and the cfg3 rank sweep (build / apply / iterations per preconditioner variant and k).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt"
NCU="ncu --clock-control none --profile-from-start off"
WIN="assemble:0:1,pchol:3000:2,syrk:0:1,potrf:10:1,trsm:5:1,pcg:5:2"

timeout 1500 python bench.py > gpurun_out/r01g_bench_default.json 2> gpurun_out/r01g_bench_default.err; stamp "default bench rc=$?"
tail -3 gpurun_out/r01g_bench_default.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r01g_bench_reference.json 2> gpurun_out/r01g_bench_reference.err; stamp "reference arm rc=$?"

timeout 300 $B > gpurun_out/r01g_head_plain.log 2>&1 && {
  stamp "head plain ok"
  MLFFPC_PROFILE=$WIN timeout 900 $NCU --metrics gpu__time_duration.sum --csv \
      --log-file gpurun_out/r01g_launches_windows.csv $B > gpurun_out/r01g_launches_stdout.log 2>&1
  stamp "launch list rc=$?"
}

timeout 900 python scripts/sweep_cfg3.py --mode matrix_free --ks 100,500,1000,2000,5000 \
    --variants cholesky,random_scores,lev_random,truncated_cholesky --tol 1e-6 --maxiter 3000 \
    > gpurun_out/r01g_sweep_cfg3.jsonl 2> gpurun_out/r01g_sweep_cfg3.err; stamp "cfg3 sweep rc=$?"
tail -3 gpurun_out/r01g_sweep_cfg3.err
wc -l gpurun_out/r01g_sweep_cfg3.jsonl
ls -la gpurun_out | tail -8
