#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
echo "== batch tests (TINY, batch engine)"; timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "batch" 2>&1 | tail -6
echo "== batch tests, kernel chain"; LNB_BATCH_ENGINE=0 timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "batch" 2>&1 | tail -3
echo "== batch bench engine"; for m in fast strict; do timeout 200 python tools/batch_bench.py $m 8 32 2>&1 | tail -1; done
echo "== batch bench chain"; for m in fast strict; do LNB_BATCH_ENGINE=0 timeout 200 python tools/batch_bench.py $m 8 16 2>&1 | tail -1; done
} > gpurun_out/r2_batch.log 2>&1
cat gpurun_out/r2_batch.log
