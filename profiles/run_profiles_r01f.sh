#!/bin/bash
# Round-1 GPU evidence, orthonormal-form preconditioner: parity, A/B of the two forms, headline bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01f_gpu_parity.log 2>&1; stamp "parity rc=$?"
tail -15 gpurun_out/r01f_gpu_parity.log
bash scripts/run_ab_numerics.sh 1e-6 "--precon-form woodbury" "--precon-form orthonormal" "--precon-form woodbury --opt syrk_chunk=2048" "--precon-form orthonormal --opt syrk_chunk=2048"
stamp "A/B done"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r01f_gpu_fullsize.log 2>&1; stamp "fullsize rc=$?"
tail -5 gpurun_out/r01f_gpu_fullsize.log
