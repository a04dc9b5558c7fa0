#!/usr/bin/env bash
# compute-sanitizer over the TINY decode / prefill / batch paths.  Run under gpurun:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# Output: gpurun_out/r02_sanitizer_{memcheck,racecheck,synccheck}.txt (copy the summaries to profiles/).
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool" > gpurun_out/r02_sanitizer_$tool.txt
  timeout 240 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_run.py >> gpurun_out/r02_sanitizer_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/r02_sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run\]|exit code" gpurun_out/r02_sanitizer_$tool.txt | tail -8
done
