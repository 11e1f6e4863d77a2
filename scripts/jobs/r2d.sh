#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or set_frame or shard" > gpurun_out/r2d_pytest_sweep.log 2>&1; echo "pytest sweep rc=$?"
tail -5 gpurun_out/r2d_pytest_sweep.log
MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_phases.so timeout 300 python scripts/gpu_sweep_timeline.py C3 C2 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2d_timeline.txt; cat gpurun_out/r2d_timeline.txt
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
for f in ['r2d_bench']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, 'ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'])
        if 'extra' in d: print(' C2', d['extra']['C2']['ms_per_step'], d['extra']['C2']['e2e_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
