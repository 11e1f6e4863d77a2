#!/bin/bash
# A/B on ONE box: 16-byte row pairs + two LDG.128 (t0) against the 16-byte patch texel + one LDG.128 (default build)
mkdir -p gpurun_out
L=$PWD/mba-vo_b200/lib
for n in 0 3 0 3; do
  MBAVO_LIBRARY=$L/libmbavo_pht$n.so timeout 300 python scripts/gpu_sweep_timeline.py C3 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2v_timeline_t$n.txt
  echo "texel $n"; awk '/pass/{printf "%s%s %.1f | ", $3,$5, $20-$15; e=$32} END{print "sweep end", e}' gpurun_out/r2v_timeline_t$n.txt
done
for v in t0 b200 t0 b200; do
  MBAVO_LIBRARY=$L/libmbavo_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2v_bench_$v.json 2> gpurun_out/r2v_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2v_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "dominant", d["roofline"].get("dominant_pass"), d["parity"] if "parity" in d else "")
except Exception as e:
    print("$v", "failed", e)
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
