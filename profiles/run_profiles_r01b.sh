#!/bin/bash
# Round-1 GPU evidence, second pass: the TMA-fed symmetric operator (symtma.cu) and the re-blocked TRSM.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
NCU="ncu --clock-control none --profile-from-start off"

timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01b_gpu_parity.log 2>&1; stamp "parity rc=$?"
tail -5 gpurun_out/r01b_gpu_parity.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r01b_gpu_fullsize.log 2>&1; stamp "fullsize rc=$?"
tail -5 gpurun_out/r01b_gpu_fullsize.log

for mode in assembled_sym assembled; do
  timeout 600 $B --tol 1e-3 --mode $mode > gpurun_out/r01b_mode_$mode.json 2> gpurun_out/r01b_mode_$mode.err; stamp "mode $mode rc=$?"
  tail -3 gpurun_out/r01b_mode_$mode.err
done

S="$B --tol 1e-2 --mode assembled_sym"
timeout 300 $S > gpurun_out/r01b_sym_plain.log 2>&1 && {
  MLFFPC_PROFILE=pcg:5:1,trsm:10:1 timeout 600 $NCU --set full --import-source on -k regex:'symv_tma|dgemm_kernel|gemv_rows' -c 8 \
      -o gpurun_out/r01b_symv_tma -f $S > gpurun_out/r01b_symv_stdout.log 2>&1
  stamp "symv_tma capture rc=$?"
}
# the headline command, as the driver runs it
timeout 1500 python bench.py > gpurun_out/r01b_bench_default.json 2> gpurun_out/r01b_bench_default.err; stamp "default bench rc=$?"
tail -3 gpurun_out/r01b_bench_default.err
ls -la gpurun_out/ | tail -12
