#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|tpp|group|boxqp|large" gpurun_out/sanitizer_$tool.txt | head -20
done
