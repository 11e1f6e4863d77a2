#!/bin/bash
# round-2 GPU job A: tests, the new C3 bench, cost-pass variants, launch list, ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.log
nproc >> gpurun_out/r2a_smi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 50 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2a_bench.err
for v in old magic u2 u8; do MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_$v.so timeout 300 python scripts/gpu_variants.py $v C2:0 C3:0 C5:0 > /dev/null; done
timeout 300 python scripts/gpu_variants.py new C2:0 C3:0 C5:0 > /dev/null
cat gpurun_out/variants.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['tag'], r['config'], 'H %.1f us (%.3f)  C %.1f us (%.3f)' % (r['us_h'], r['frac_h'], r['us_c'], r['frac_c']))"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:track_kernel -c 4 -f -o gpurun_out/r2a_c3l0 python scripts/ncu_target.py C3 0 2 > gpurun_out/r2a_ncu_full.log 2>&1; echo "ncu full rc=$?"
