#!/bin/bash
# patch texel (default build) with the explicit blend: A/B against the row-pair build, unroll variants, ncu capture, full GPU suite
mkdir -p gpurun_out
L=$PWD/mba-vo_b200/lib
for n in 0 3; do
  MBAVO_LIBRARY=$L/libmbavo_pht$n.so timeout 300 python scripts/gpu_sweep_timeline.py C3 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2w_timeline_t$n.txt
  echo "texel $n"; awk '/pass/{printf "%s%s %.1f | ", $3,$5, $20-$15; e=$32} END{print "sweep end", e}' gpurun_out/r2w_timeline_t$n.txt
done
for v in t0 b200 u1 u3 u4 b200; do
  MBAVO_LIBRARY=$L/libmbavo_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2w_bench_$v.json 2> gpurun_out/r2w_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2w_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"]["value"], "frac", round(d["roofline"]["frac"],4), "L0H us", d["roofline"]["dominant_pass"]["us"], [ (p["pass"], round(p["us"],1)) for p in d["roofline"]["passes"]])
except Exception as e:
    print("$v", "failed", e)
PY
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -f -o gpurun_out/r2w_sweep_c3 python scripts/ncu_sweep_target.py C3 4 > gpurun_out/r2w_ncu_full.log 2>&1; echo "ncu full rc=$?"
