#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 50 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
PY
