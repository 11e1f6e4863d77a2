#!/bin/bash
mkdir -p gpurun_out
MBAVO_LIBRARY=$PWD/mba-vo_b200/lib/libmbavo_phases.so timeout 300 python scripts/gpu_sweep_timeline.py C3 C2 2>&1 | grep -A10 "persistent launch" | cut -c1-330 > gpurun_out/r2q_timeline.txt; awk '/pass/{print $1,$2,$3,$4,$5, $12,$13, $14,$15,$16,$17, $18,$19,$20,$21,$22,$23,$24,$25,$26,$27,$28,$29,$30,$31,$32,$33,$34,$35,$36,$37,$38,$39,$40,$41,$42,$43}' gpurun_out/r2q_timeline.txt | cut -c1-280
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or set_frame or shard or baseline" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
