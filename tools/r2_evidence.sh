#!/usr/bin/env bash
# Round-2 evidence call: sanitizer on both decode paths, ncu capture of the engine, bench lines of every config at 1 GPU.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_evidence.log
{
  echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
  echo "== bench.py (default config)"; timeout 600 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo rc=$?; head -c 600 gpurun_out/r2b_bench_n1.json; echo
  echo "== bench.py --config batch8"; timeout 600 python bench.py --config batch8 --steps 2 --parity-tokens 8 > gpurun_out/r2b_bench_batch8_n1.json 2> gpurun_out/r2b_bench_batch8_n1.err; echo rc=$?; head -c 900 gpurun_out/r2b_bench_batch8_n1.json; echo
  echo "== bench.py --config prefill2048"; timeout 600 python bench.py --config prefill2048 --steps 3 > gpurun_out/r2b_bench_prefill_n1.json 2> gpurun_out/r2b_bench_prefill_n1.err; echo rc=$?; head -c 900 gpurun_out/r2b_bench_prefill_n1.json; echo
  echo "== ncu launch list: strict decode, engine (prefill chain + 1 engine launch of 3 steps)"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_strict_engine.csv python tools/prof_engine.py strict 3 > gpurun_out/r2_prof_engine.log 2>&1
  python tools/ncu_summarize.py gpurun_out/r2_launches_strict_engine.csv 2>&1 | head -14
  echo "== ncu --set full: the engine kernel (strict, 2 steps)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_engine -c 1 -o gpurun_out/r2_engine_strict python tools/prof_engine.py strict 2 > gpurun_out/r2_prof_engine_full.log 2>&1; echo rc=$?
  tail -3 gpurun_out/r2_prof_engine_full.log
} > "$OUT" 2>&1
bash tools/sanitize.sh >> "$OUT" 2>&1
tail -70 "$OUT"
