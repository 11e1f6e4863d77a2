#!/bin/bash
# Serial section of a Hessian pass: wide single-round sum of the partials (ws), knots / candidate through shared memory +
# precomputed sample positions + series exp (ff), both (wsff) against the default build, on ONE box; GPU suite with the winner.
mkdir -p gpurun_out
L=$PWD/mba-vo_b200/lib
VARIANTS="${VARIANTS:-b200 ws ff wsff b200_again}"
for v in $VARIANTS; do
  lib=$L/libmbavo_${v%_again}.so
  MBAVO_LIBRARY=$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${TAG:-r2y}_bench_$v.json 2> gpurun_out/${TAG:-r2y}_bench_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG:-r2y}_bench_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4), {k: round(p["us"],1) for k,p in d["roofline"]["passes"].items()}, "parity", d["parity"]["cost_rel_max"], d["parity"]["first_lm_step_rel_max"])
except Exception as e:
    print("$v", "failed", e)
PY
done
BEST=$(VARIANTS="$VARIANTS" TAG=${TAG:-r2y} python - <<'PY'
import json, os
res = {}
for v in os.environ["VARIANTS"].split():
    try:
        res[v] = json.loads(open(f"gpurun_out/{os.environ['TAG']}_bench_{v}.json").read().strip().splitlines()[-1])["ms_per_step"]
    except Exception:
        pass
base = min(res.get("b200", 9e9), res.get("b200_again", 9e9))
best, bt = "b200", None
for v, t in res.items():
    if not v.startswith("b200") and t < 0.992 * base and (bt is None or t < bt):
        best, bt = v, t
print(best)
PY
)
echo "winner: $BEST"
if [ "$BEST" != "b200" ]; then
  export MBAVO_LIBRARY=$L/libmbavo_$BEST.so
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG:-r2y}_pytest_$BEST.log 2>&1; echo "pytest($BEST) rc=$?"; tail -3 gpurun_out/${TAG:-r2y}_pytest_$BEST.log
fi
