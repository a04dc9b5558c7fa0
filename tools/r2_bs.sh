#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
echo "== batch tests"; timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "batch" 2>&1 | tail -5
echo "== batch bench strict"; timeout 200 python tools/batch_bench.py strict 8 32 2>&1 | tail -1
echo "== profile"; timeout 200 python tools/batch_prof.py strict 2>&1 | tail -12
} > gpurun_out/r2_bs.log 2>&1
cat gpurun_out/r2_bs.log
