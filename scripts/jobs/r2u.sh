#!/bin/bash
# A/B of the keyframe texel format on ONE box: 0 = 16-byte row pairs (two LDG.128), 1 = 32-byte texel (one LDG.256), 2 = 32-byte texel (two LDG.128)
mkdir -p gpurun_out
L=$PWD/mba-vo_b200/lib
for n in 0 1 2; do
  MBAVO_LIBRARY=$L/libmbavo_pht$n.so timeout 300 python scripts/gpu_sweep_timeline.py C3 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2u_timeline_t$n.txt
  echo "texel $n"; awk '/pass/{printf "%s%s %.1f | ", $3,$5, $20-$15; e=$32} END{print "sweep end", e}' gpurun_out/r2u_timeline_t$n.txt
done
for v in t0 b200 t2 t0 b200 t2; do
  MBAVO_LIBRARY=$L/libmbavo_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2u_bench_$v.json 2> gpurun_out/r2u_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2u_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "dominant", d["roofline"].get("dominant_pass"))
except Exception as e:
    print("$v", "failed", e)
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
