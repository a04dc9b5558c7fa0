#!/usr/bin/env bash
# Round-2 GPU call #2: the persistent decode engine.   gpurun --timeout 1200 -- 'bash tools/r2_call2.sh'
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call2.log
{
  echo "== engine tests (TINY, both modes, vs chain and oracle)"; timeout 300 python -m pytest tests/test_gpu_engine.py -q -x 2>&1 | tail -25
  echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -25
  echo "== kbench engine"; timeout 200 python tools/kbench.py fast,strict 2>&1 | tail -4
  echo "== kbench chain"; LNB_ENGINE=0 timeout 200 python tools/kbench.py fast,strict 2>&1 | tail -4
} > "$OUT" 2>&1
tail -70 "$OUT"
