#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
echo "== sdpa_tc2 vs oracle"; timeout 150 python tools/sdpa_tc_check.py 33 70 128 200 513 2>&1 | tail -6
echo "== prefill2048 bench, LNB_SDPA_TC=2"; LNB_SDPA_TC=2 timeout 300 python bench.py --config prefill2048 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bench_prefill2048_e.json | cut -c1-330
echo "== prefill2048 bench, default (sdpa_tc_kernel)"; timeout 300 python bench.py --config prefill2048 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-200
echo "== model tests, LNB_SDPA_TC=2"; LNB_SDPA_TC=2 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k attention 2>&1 | tail -3
} > gpurun_out/r2_sdpa2.log 2>&1
cat gpurun_out/r2_sdpa2.log
