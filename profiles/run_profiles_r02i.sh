#!/bin/bash
# Round 2 (1 GPU): full suite with the second-generation pair kernel, the probe-based 'auto' operator choice and smoke();
# pair-kernel A/B at cfg2 and on the cfg5 per-rank slice; then the ncu evidence of run_profiles_r02g.sh.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02i_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02i_gpu_tests.log | cut -c1-300 | head -40
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02i_smoke.log 2>&1; stamp "smoke rc=$?"
tail -6 gpurun_out/r02i_smoke.log
for PK in 2 1; do
  timeout 300 python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --pairs-kernel $PK > gpurun_out/r02i_mf_cfg2_pk$PK.json 2>&1
  timeout 300 python scripts/matvec_free_bench.py --pairs-kernel $PK > gpurun_out/r02i_mf_cfg5_slice_pk$PK.json 2>&1
  tail -1 gpurun_out/r02i_mf_cfg2_pk$PK.json | cut -c1-400; tail -1 gpurun_out/r02i_mf_cfg5_slice_pk$PK.json | cut -c1-400
done
stamp "pair kernel A/B"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02i_bench_default.json 2> gpurun_out/r02i_bench_default.err; stamp "bench default rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02i_bench_default.json'))
    print('value', d['value'], 'e2e', d['e2e']['value'], 'alt', d.get('alt'))
    for s in d['phases']['per_step']: print('  ', s)
except Exception as e:
    print('parse failed', e)
PY
bash profiles/run_profiles_r02g.sh
