#!/bin/bash
# compute-sanitizer over tools/sanitize_cases.py; summaries land in gpurun_out/sanitizer_<tool>.log
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
