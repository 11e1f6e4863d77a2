#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 1 -f -o gpurun_out/r2t_sweep_c3 python scripts/ncu_sweep_target.py C3 4 > gpurun_out/r2t_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -2 gpurun_out/r2t_ncu_full.log
