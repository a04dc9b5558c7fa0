#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
echo "== batch tests (TINY, tensor-core FAST + STRICT)"; timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "batch" 2>&1 | tail -8 | tee /tmp/t.log
if grep -q failed /tmp/t.log; then
  echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_model.py -q -x -k "batched_decode_equals_independent_contexts and fast" 2>&1 | grep -v "^$" | head -60
else
echo "== batch bench"; for m in fast; do timeout 200 python tools/batch_bench.py $m 8 32 2>&1 | tail -1; done
echo "== profile"; timeout 200 python tools/batch_prof.py fast 2>&1 | tail -12
fi
} > gpurun_out/r2_tc.log 2>&1
cat gpurun_out/r2_tc.log
