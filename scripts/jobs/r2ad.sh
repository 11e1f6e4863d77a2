#!/bin/bash
# Final build of round 2: GPU suite, smoke, full bench, reference arm, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ad_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ad_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2ad_smoke.log
timeout 900 python bench.py > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2ad_bench_ref.json 2> gpurun_out/r2ad_bench_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ad_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ad_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ad_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches', d['gpu_launches'], 'steps', d['steps'])
r=d['roofline']; print({k:(round(v['us'],1), round(v['frac'],3)) for k,v in r['passes'].items()}, r['kernel_ms'], r['frac'], r['kernel_share_of_step'])
print('standalone', r['standalone_level0_hessian_frac'])
print('gpu_baseline', d['gpu_baseline']['ms_per_step'], 'cpu', d['cpu_baseline'])
print('parity', d['parity']['cost_rel_max'], d['parity']['first_lm_step_rel_max'])
print('C2', json.dumps(d['extra']['C2'])[:1200])
r=json.loads(open('gpurun_out/r2ad_bench_ref.json').read().strip().splitlines()[-1]); print('ref arm', r['ms_per_step'], r['value'], r['cpu_baseline']['cores'])
PY
