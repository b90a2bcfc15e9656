#!/bin/bash
# Round-1 GPU evidence (run under gpurun, one GPU).  Numbers printed by runs under ncu are never bench values.
# ncu sees only the library's profiler windows (MLFFPC_PROFILE -> cudaProfilerStart/Stop around chosen
# pivot steps / CG iterations / factor phases; --profile-from-start off), so a 38 000-launch solve yields a
# ~40-launch list instead of an hour of serialised replays.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
NCU="ncu --clock-control none --profile-from-start off"
WIN="assemble:0:1,pchol:3000:2,syrk:0:1,potrf:10:1,trsm:20:1,pcg:5:2"

# (0) parity first
python -m pytest tests -m gpu -x -q > gpurun_out/r01_gpu_tests.log 2>&1; stamp "gpu tests rc=$?"
tail -3 gpurun_out/r01_gpu_tests.log

# (1) fp64 GEMM rates (cuBLAS vs our DMMA kernel)
python scripts/fp64_peak.py > gpurun_out/r01_fp64_peak.json 2> gpurun_out/r01_fp64_peak.err; stamp "fp64 peak rc=$?"

# (2) operator comparison at the headline size, short CG (tol 1e-3): plain GEMV, symmetric, matrix-free
for mode in assembled assembled_sym matrix_free; do
  $B --tol 1e-3 --mode $mode > gpurun_out/r01_mode_$mode.json 2> gpurun_out/r01_mode_$mode.err; stamp "mode $mode rc=$?"
done

# (3) launch list: the headline command (tol 1e-6), profiler windows only
$B > gpurun_out/r01_head_plain.log 2>&1 && {
  stamp "head plain ok"
  MLFFPC_PROFILE=$WIN timeout 900 $NCU --metrics gpu__time_duration.sum --csv \
      --log-file gpurun_out/r01_launches_windows.csv $B > gpurun_out/r01_launches_stdout.log 2>&1
  stamp "launch list rc=$?"
}

# (4) full captures of the hot kernels inside the same windows (short CG so the run is brief)
P="$B --tol 1e-2"
$P > gpurun_out/r01_prof_plain.log 2>&1 && {
  stamp "prof plain ok"
  MLFFPC_PROFILE=$WIN timeout 900 $NCU --set full --import-source on \
      -k regex:'gemv_rows_kernel|tgemv_cols_kernel|pchol_update_kernel|assemble_block_kernel|dgemm_kernel' -c 14 \
      -o gpurun_out/r01_hot -f $P > gpurun_out/r01_hot_stdout.log 2>&1
  stamp "hot capture rc=$?"
}
S="$B --tol 1e-2 --mode assembled_sym"
$S > gpurun_out/r01_sym_plain.log 2>&1 && {
  MLFFPC_PROFILE=pcg:5:1 timeout 600 $NCU --set full --import-source on -k regex:'symv_' -c 2 \
      -o gpurun_out/r01_symv -f $S > gpurun_out/r01_symv_stdout.log 2>&1
  stamp "symv capture rc=$?"
}
ls -la gpurun_out/
