#!/bin/bash
# Third cut of the serial sections: rcp.approx.f64 seed of the pivot reciprocals (r64), + selector / commit inputs prefetched under the
# sum (f3), against the default build on one box; GPU suite with the winner.  (Last GPU minutes of the round: tight time-outs.)
mkdir -p gpurun_out
L=$PWD/mba-vo_b200/lib
for v in b200 r64 f3; do
  MBAVO_LIBRARY=$L/libmbavo_$v.so timeout 60 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2ah_bench_$v.json 2> gpurun_out/r2ah_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2ah_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "kernel", round(d["roofline"]["kernel_ms"],4), {k: round(p["us"],1) for k,p in d["roofline"]["passes"].items()})
except Exception as e:
    print("$v", "failed", e)
PY
done
BEST=$(python - <<'PY'
import json
res = {}
for v in ("b200", "r64", "f3"):
    try:
        res[v] = json.loads(open(f"gpurun_out/r2ah_bench_{v}.json").read().strip().splitlines()[-1])["roofline"]["kernel_ms"]
    except Exception:
        pass
best, bt = "b200", res.get("b200", 9e9) * 0.997
for v in ("r64", "f3"):
    if v in res and res[v] < bt:
        best, bt = v, res[v]
print(best)
PY
)
echo "winner: $BEST"
if [ "$BEST" != "b200" ]; then
  MBAVO_LIBRARY=$L/libmbavo_$BEST.so timeout 130 python -m pytest tests -m gpu -x -q > gpurun_out/r2ah_pytest_$BEST.log 2>&1; echo "pytest($BEST) rc=$?"; tail -3 gpurun_out/r2ah_pytest_$BEST.log
fi
