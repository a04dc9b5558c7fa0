#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
echo "== ops + model tests"; timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -q -x 2>&1 | tail -3
echo "== prefill2048 bench"; timeout 300 python bench.py --config prefill2048 --steps 5 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/r02z_bench_prefill2048_3p.json | cut -c1-330
} > gpurun_out/r2_pf.log 2>&1
cat gpurun_out/r2_pf.log
