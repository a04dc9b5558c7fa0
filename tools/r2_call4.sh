#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_call4.log
{
  echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
  echo "== racecheck, engine path"; LNB_ENGINE=1 LNB_P2P_TIMEOUT_MS=0 timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_run.py > gpurun_out/r02_sanitizer_racecheck_engine1.txt 2>&1; grep -E "RACECHECK SUMMARY|sanitize_run\]" gpurun_out/r02_sanitizer_racecheck_engine1.txt | tail -4
  echo "== bench.py --config prefill2048 --parity"; timeout 900 python bench.py --config prefill2048 --steps 3 --parity > gpurun_out/r2b_bench_prefill_parity.json 2> gpurun_out/r2b_bench_prefill_parity.err; echo rc=$?; tail -c 1500 gpurun_out/r2b_bench_prefill_parity.json; tail -2 gpurun_out/r2b_bench_prefill_parity.err
} > "$OUT" 2>&1
tail -30 "$OUT"
