#!/bin/bash
# MBAVO_UPLOAD_DEFER_POINTS: the fine levels' point copies issued behind the launch of the persistent sweep. Test first, then A/B of the end-to-end leg.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "set_frame" > gpurun_out/r2aa_pytest_setframe.log 2>&1; echo "pytest set_frame rc=$?"; tail -3 gpurun_out/r2aa_pytest_setframe.log
for v in 0 1 0 1; do
  MBAVO_BENCH_DEFER=$v timeout 300 python bench.py --steps 50 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2aa_bench_defer$v.json 2> gpurun_out/r2aa_bench_defer$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2aa_bench_defer$v.json").read().strip().splitlines()[-1])
    print("defer=$v", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4))
except Exception as e:
    print("defer=$v", "failed", e)
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2aa_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aa_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2aa_smoke.log
