#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call3.log
{
  echo "== engine tests"; timeout 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_8b.py tests/test_gpu_model.py tests/test_gpu_ops.py -q -x 2>&1 | tail -8
  echo "== decode"; timeout 400 python tools/engine_sweep.py 0 2>&1 | tail -3
  echo "== engine profile"; LNB_ENGINE=1 timeout 300 python tools/engine_prof.py ${MODES:-strict} 2>&1 | tail -40
} > "$OUT" 2>&1
tail -60 "$OUT"
