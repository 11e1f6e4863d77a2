#!/bin/bash
# round-2 GPU job B: persistent sweep kernel — tests first (bounded), then bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or set_frame or shard" > gpurun_out/r2b_pytest_sweep.log 2>&1; echo "pytest sweep rc=$?"
tail -15 gpurun_out/r2b_pytest_sweep.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.err
MBAVO_NO_PERSISTENT=1 timeout 600 python bench.py --steps 50 --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_nopersist.json 2> gpurun_out/r2b_bench_np.err; echo "bench np rc=$?"
python - <<'PY'
import json
for f in ['r2b_bench','r2b_bench_nopersist']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, 'ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'])
        if 'extra' in d: print(' C2', d['extra']['C2']['ms_per_step'], d['extra']['C2']['e2e_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
