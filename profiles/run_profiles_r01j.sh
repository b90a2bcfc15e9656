#!/bin/bash
# split-K matrix-free contraction: parity + timings (one GPU)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r01j_gpu_parity.log 2>&1; echo "parity rc=$?"
tail -4 gpurun_out/r01j_gpu_parity.log
timeout 300 python scripts/matvec_free_bench.py --reps 5 | tee gpurun_out/r01j_mf_cfg5_slice.json
timeout 300 python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --reps 5 | tee gpurun_out/r01j_mf_cfg2.json
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --mode matrix_free > gpurun_out/r01j_bench_matrix_free.json 2> gpurun_out/r01j_bench_matrix_free.err; echo "bench mf rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r01j_bench_matrix_free.json')); print('matrix_free cfg2: value', d['value'], d['phases'])"
