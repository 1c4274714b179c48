#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize.py > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "$tool rc $?"
  grep -E "ERROR SUMMARY|sanitize workload ok|RACECHECK SUMMARY|hazard" gpurun_out/r2_sanitizer_$tool.log | head -5
done
