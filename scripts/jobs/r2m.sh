#!/bin/bash
mkdir -p gpurun_out
MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_phases.so timeout 300 python scripts/gpu_sweep_timeline.py C3 C2 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2m_timeline.txt; awk '{print $1,$2,$3,$4,$5, $12,$13, $14,$15,$16,$17, $18,$19,$20,$21,$22,$23,$24,$25,$26,$27,$28,$29,$30,$31,$32,$33,$34,$35,$36,$37,$38,$39,$40,$41,$42,$43}' gpurun_out/r2m_timeline.txt | cut -c1-280
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 50 --no-cpu-baseline --no-extras > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
r=d['roofline']; print({k:round(v['us'],1) for k,v in r['passes'].items()}, r['kernel_ms'], r['frac'])
PY
