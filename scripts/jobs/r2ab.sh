#!/bin/bash
# prepared C calls in the bench's timed loop (argument marshalling once), deferred point copies on / off
mkdir -p gpurun_out
for v in 1 0 1 0; do
  MBAVO_BENCH_DEFER=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2ab_bench_defer$v.json 2> gpurun_out/r2ab_bench_defer$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2ab_bench_defer$v.json").read().strip().splitlines()[-1])
    print("defer=$v", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4))
except Exception as e:
    print("defer=$v", "failed", e)
PY
done
timeout 900 python bench.py --steps 100 > gpurun_out/r2ab_bench_full.json 2> gpurun_out/r2ab_bench_full.err; echo "bench full rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ab_bench_full.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], 'launches', d['gpu_launches'])
print('C2', json.dumps(d['extra']['C2'])[:400])
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ab_pytest.log
