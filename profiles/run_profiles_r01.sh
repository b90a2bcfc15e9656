#!/bin/bash
# Round-1 ncu evidence (run under gpurun, one GPU).  Numbers printed by runs under ncu are never bench values.
# Every ncu run is bounded (-c) and wrapped in `timeout`: an unbounded launch list of a 38 700-launch step
# costs ~80 ms per launch under ncu and never finishes.
set -u
mkdir -p gpurun_out
HEAD="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"            # the headline command (cfg2, tol 1e-6)
PROF="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --tol 1e-3"   # same kernels, shorter CG
SYM="$PROF --mode assembled_sym"
NCU="ncu --clock-control none"

# (1) launch lists: two 300-launch windows of the headline step (pivoted-Cholesky phase, PCG phase)
$HEAD > gpurun_out/r01_head_plain.log 2>&1 && {
  timeout 400 $NCU --metrics gpu__time_duration.sum -s 12000 -c 300 --csv \
      --log-file gpurun_out/r01_launches_pchol_window.csv $HEAD > /dev/null 2>&1; echo "list pchol rc=$?"
  timeout 400 $NCU --metrics gpu__time_duration.sum -s 26000 -c 320 --csv \
      --log-file gpurun_out/r01_launches_pcg_window.csv $HEAD > /dev/null 2>&1; echo "list pcg rc=$?"
}

# (2) full captures of the hot kernels at the headline size (n = 108 000)
$PROF > gpurun_out/r01_prof_plain.log 2>&1 && {
  timeout 300 $NCU --set full --import-source on -k regex:gemv_rows_kernel -s 2 -c 2 \
      -o gpurun_out/r01_gemv -f $PROF > /dev/null 2>&1; echo "gemv rc=$?"
  timeout 300 $NCU --set full --import-source on -k regex:pchol_update_kernel -s 3000 -c 2 \
      -o gpurun_out/r01_pchol_update -f $PROF > /dev/null 2>&1; echo "pchol rc=$?"
  timeout 300 $NCU --set full --import-source on -k regex:tgemv_cols_kernel -s 2 -c 1 \
      -o gpurun_out/r01_tgemv -f $PROF > /dev/null 2>&1; echo "tgemv rc=$?"
  timeout 300 $NCU --set full --import-source on -k regex:dgemm_kernel -s 1 -c 1 \
      -o gpurun_out/r01_dgemm -f $PROF > /dev/null 2>&1; echo "dgemm rc=$?"
  timeout 300 $NCU --set full --import-source on -k regex:assemble_block_kernel -s 1 -c 1 \
      -o gpurun_out/r01_assemble -f $PROF > /dev/null 2>&1; echo "assemble rc=$?"
}
$SYM > gpurun_out/r01_sym_plain.log 2>&1 && {
  timeout 300 $NCU --set full --import-source on -k regex:symv_strip_kernel -s 2 -c 1 \
      -o gpurun_out/r01_symv -f $SYM > /dev/null 2>&1; echo "symv rc=$?"
}
ls -la gpurun_out/
