#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r2_sanitizer.txt
: > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool python scripts/sanitize_target.py" >> $out
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py >> $out 2>&1
  echo "rc=$?" >> $out
done
grep -n "ERROR SUMMARY\|== compute\|rc=\|Error\|error" $out | head -40
