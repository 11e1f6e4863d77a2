#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2r_pytest.log
timeout 900 python bench.py --steps 100 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2r_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2r_bench_ref.json; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2r_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2r_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches', d['gpu_launches'])
r=d['roofline']; print({k:(round(v['us'],1), round(v['frac'],3)) for k,v in r['passes'].items()}, r['kernel_ms'], r['frac'], r['kernel_share_of_step'])
print('standalone', r['standalone_level0_hessian_frac'])
print('gpu_baseline', d['gpu_baseline']['ms_per_step'], 'cpu', d['cpu_baseline'])
print('parity', d['parity']['cost_rel_max'], d['parity']['first_lm_step_rel_max'])
print('C2', json.dumps(d['extra']['C2'])[:900])
r=json.load(open('gpurun_out/r2r_bench_ref.json')); print('ref arm', r['ms_per_step'], r['value'], r['cpu_baseline']['cores'])
PY
