#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 50 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -f -o gpurun_out/r2f_sweep_c3 python scripts/ncu_sweep_target.py C3 4 > gpurun_out/r2f_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/r2f_ncu_full.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
r=d['roofline']; print('roofline frac', r['frac'], 'kernel_ms', r['kernel_ms'], 'share', r['kernel_share_of_step'])
print({k:(round(v['us'],1), round(v['frac'],3)) for k,v in r['passes'].items()})
print('standalone L0H frac', r['standalone_level0_hessian_frac'])
print('gpu_baseline', d['gpu_baseline']['ms_per_step'], 'cpu', d['cpu_baseline']['ms_per_step'])
print('C2', d['extra']['C2']['ms_per_step'], d['extra']['C2']['e2e_ms_per_step'], d['extra']['C2']['lm_to_convergence'])
PY
