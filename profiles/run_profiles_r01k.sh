#!/bin/bash
# row-walking assembly kernel: parity + A/B against the first-generation kernel (one GPU)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01k_gpu_parity.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/r01k_gpu_parity.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --tol 1e-2"
for cfg in "--mode assembled_sym" "--mode assembled_sym --opt assemble_legacy=1" "--mode assembled" "--mode assembled --opt assemble_legacy=1"; do
  timeout 600 $B $cfg > gpurun_out/r01k_tmp.json 2> gpurun_out/r01k_tmp.err
  python -c "
import json; d=json.load(open('gpurun_out/r01k_tmp.json')); print('[$cfg] assemble_s %.4f  value %.2f iters %d' % (d['phases']['assemble_s'], d['value'], d['phases']['cg_iters']))"
  tail -2 gpurun_out/r01k_tmp.err
done
S="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --tol 1e-2 --mode assembled"
timeout 300 $S > gpurun_out/r01k_plain.log 2>&1 && {
  MLFFPC_PROFILE=assemble:0:1 timeout 600 ncu --clock-control none --profile-from-start off --set full --import-source on \
      -k regex:'assemble_rows' -c 1 -o gpurun_out/r01k_assemble -f $S > gpurun_out/r01k_ncu_stdout.log 2>&1
  echo "assemble capture rc=$?"
}
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r01k_gpu_fullsize.log 2>&1; echo "fullsize rc=$?"
tail -3 gpurun_out/r01k_gpu_fullsize.log
