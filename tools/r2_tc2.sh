#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
echo "== full gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== prefill2048 bench"; timeout 900 python bench.py --config prefill2048 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bench_prefill2048_b.json
echo "== batch8 bench fast"; timeout 600 python bench.py --config batch8 --acc fast --steps 32 --warmup 4 2>&1 | tail -1 | tee gpurun_out/r2_bench_batch8_fast.json
} > gpurun_out/r2_tc2.log 2>&1
cat gpurun_out/r2_tc2.log
