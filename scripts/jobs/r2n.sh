#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2n_pytest.log
for e in 0 1; do
echo "== MBAVO_NO_SMALL_SPLIT=$e"
MBAVO_NO_SMALL_SPLIT=$e MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_phases.so timeout 300 python scripts/gpu_sweep_timeline.py C3 C2 2>&1 | grep -A10 "persistent launch" | awk '/pass/{print $1,$2,$3,$4,$5, $12,$13, $14,$15,$16,$17, $18,$19,$20,$21,$22,$23,$24,$25,$26,$27,$28,$29,$30,$31,$32}' | cut -c1-200
done
timeout 900 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
r=d['roofline']; print({k:round(v['us'],1) for k,v in r['passes'].items()}, r['kernel_ms'], r['frac'])
print('C2', d['extra']['C2']['ms_per_step'], d['extra']['C2']['e2e_ms_per_step'], d['extra']['C2'].get('shim'))
PY
