#!/bin/bash
# compute-sanitizer over the final build's persistent sweep (quick target): memcheck, racecheck, initcheck
mkdir -p gpurun_out
: > gpurun_out/r2ag_sanitizer.txt
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool python scripts/sanitize_target.py --quick" >> gpurun_out/r2ag_sanitizer.txt
  timeout 70 compute-sanitizer --tool $tool python scripts/sanitize_target.py --quick >> gpurun_out/r2ag_sanitizer.txt 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/r2ag_sanitizer.txt
done
grep -c "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/r2ag_sanitizer.txt; grep "SUMMARY\| ok " gpurun_out/r2ag_sanitizer.txt | cut -c1-160
