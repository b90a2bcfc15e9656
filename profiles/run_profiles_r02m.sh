#!/bin/bash
# Round 2 (1 GPU): suite after the pair-stage rework (lean sqrt / exp epilogue, 128 x 32 tiles with two CTAs per SM),
# A/B of the three pair kernels at cfg2 and on the cfg5 slice, ncu of the new kernel, headline bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02m_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02m_gpu_tests.log | cut -c1-300 | head -40
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02m_smoke.log 2>&1; stamp "smoke rc=$?"
tail -5 gpurun_out/r02m_smoke.log
for PK in 1 2 3; do
  timeout 300 python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --pairs-kernel $PK > gpurun_out/r02m_mf_cfg2_pk$PK.json 2>&1
  timeout 300 python scripts/matvec_free_bench.py --pairs-kernel $PK > gpurun_out/r02m_mf_cfg5_slice_pk$PK.json 2>&1
  echo "pairs_kernel=$PK"; tail -1 gpurun_out/r02m_mf_cfg2_pk$PK.json | cut -c1-330; tail -1 gpurun_out/r02m_mf_cfg5_slice_pk$PK.json | cut -c1-330
done
stamp "pair kernel A/B"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02m_bench_default.json 2> gpurun_out/r02m_bench_default.err; stamp "bench default rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02m_bench_default.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'alt', d.get('alt'))
    for s in d['phases']['per_step']: print('  ', s)
except Exception as e:
    print('parse failed', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mv_pairs2_kernel -s 2 -c 1 -o gpurun_out/r02m_pairs3 -f \
    python scripts/matvec_free_bench.py --reps 1 --pairs-kernel 3 > gpurun_out/r02m_ncu_pairs3.log 2>&1; stamp "ncu pairs3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mv_pairs_kernel -s 2 -c 1 -o gpurun_out/r02m_pairs1 -f \
    python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --reps 1 --pairs-kernel 1 > gpurun_out/r02m_ncu_pairs1.log 2>&1; stamp "ncu pairs1 (cfg2) rc=$?"
