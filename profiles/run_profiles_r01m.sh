#!/bin/bash
# row-walking assembly kernel, third pass (product tables): parity + timing (one GPU)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01m_gpu_parity.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/r01m_gpu_parity.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-alt --tol 1e-2"
for cfg in "--mode assembled_sym" "--mode assembled"; do
  timeout 600 $B $cfg > gpurun_out/r01m_tmp.json 2> gpurun_out/r01m_tmp.err
  python -c "
import json; d=json.load(open('gpurun_out/r01m_tmp.json')); print('[$cfg] assemble_s %.4f  value %.2f iters %d' % (d['phases']['assemble_s'], d['value'], d['phases']['cg_iters']))"
  tail -2 gpurun_out/r01m_tmp.err
done
S="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-alt --tol 1e-2 --mode assembled"
timeout 300 $S > gpurun_out/r01m_plain.log 2>&1 && {
  MLFFPC_PROFILE=assemble:0:1 timeout 600 ncu --clock-control none --profile-from-start off --set full --import-source on \
      -k regex:'assemble_rows' -c 1 -o gpurun_out/r01m_assemble -f $S > gpurun_out/r01m_ncu_stdout.log 2>&1
  echo "assemble capture rc=$?"
}
