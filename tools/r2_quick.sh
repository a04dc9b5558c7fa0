#!/usr/bin/env bash
mkdir -p gpurun_out
{
echo "== tests"; timeout 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_model.py -q -x 2>&1 | tail -3
echo "== engine (LNB_ENGINE=1) decode"; LNB_ENGINE=1 timeout 300 python tools/engine_sweep.py 0 2>&1 | tail -2
echo "== engine kbench"; LNB_ENGINE=1 timeout 300 python tools/kbench.py fast 2>&1 | grep -v "graph decode" | tail -3
bash tools/r2_race.sh 2>&1 | grep -E "^=== |RACECHECK SUMMARY"
} > gpurun_out/r2_quick.log 2>&1
cat gpurun_out/r2_quick.log
