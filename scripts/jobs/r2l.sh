#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in phOLD phases; do
  echo "== $v"
  MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_$v.so timeout 300 python scripts/gpu_sweep_timeline.py C3 2>&1 | grep -A10 "persistent launch" | awk '/pass/{ if ($2 % 2 == 1) printf "%s%s L%s: load->batches %.1f  total(pass end) %s | ", $3, "", $5, $17-$10, $32 }' | cut -c1-400; echo
done; done
timeout 900 python bench.py --steps 50 --no-cpu-baseline --no-extras > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
r=d['roofline']; print({k:round(v['us'],1) for k,v in r['passes'].items()}, r['kernel_ms'], r['frac'])
PY
