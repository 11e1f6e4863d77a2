#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2k_pytest.log
MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_phases.so timeout 300 python scripts/gpu_sweep_timeline.py C3 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2k_timeline.txt; cat gpurun_out/r2k_timeline.txt
timeout 900 python bench.py --steps 50 --no-cpu-baseline --no-extras > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
r=d['roofline']; print({k:round(v['us'],1) for k,v in r['passes'].items()}, r['kernel_ms'], r['frac'])
PY
