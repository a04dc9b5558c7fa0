#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call3.log
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.limit,temperature.gpu --format=csv,noheader
  echo "== chain kbench (reference on this box)"; LNB_ENGINE=0 timeout 200 python tools/kbench.py fast 2>&1 | tail -2
  echo "== engine kbench PF=0 (nested-loop producer)"; LNB_ENGINE_PF_KB=0 timeout 200 python tools/kbench.py fast,strict 2>&1 | tail -4
  echo "== engine kbench PF=256 (iterator producer + L2 prefetch)"; LNB_ENGINE_PF_KB=256 timeout 200 python tools/kbench.py fast,strict 2>&1 | tail -4
  echo "== engine kbench PF=1 (iterator producer, ~no prefetch)"; LNB_ENGINE_PF_KB=1 timeout 200 python tools/kbench.py fast 2>&1 | tail -2
} > "$OUT" 2>&1
tail -60 "$OUT"
