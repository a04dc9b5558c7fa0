#!/usr/bin/env bash
mkdir -p gpurun_out
for m in strict fast; do
  NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=6 timeout 200 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 6 --kernel-name kns=decode_engine python tools/race_min.py $m 1 > gpurun_out/race_$m.txt 2>&1
  echo "=== $m"; grep -v "^=========\s*$" gpurun_out/race_$m.txt | head -70 | cut -c1-230
done
