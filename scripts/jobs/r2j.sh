#!/bin/bash
mkdir -p gpurun_out
for cfg in C3:0 C5:0; do
  tag=$(echo $cfg | tr ':' 'l' | tr 'A-Z' 'a-z')
  TMA_BOX=48x32 TMA_REPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"track_cost_tma_kernel|track_kernel" -s 3 -c 4 -f -o gpurun_out/r2j_tma_$tag python scripts/gpu_tma_variant.py $cfg > gpurun_out/r2j_ncu_$tag.log 2>&1; echo "ncu $cfg rc=$?"
done
