#!/usr/bin/env bash
# after the pipelined prompt attention became the default: tests, sanitizer on the attention / prefill paths, config 3 bench + launch list
set -u
mkdir -p gpurun_out
O=gpurun_out
{
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench prefill2048"; timeout 300 python bench.py --config prefill2048 2>/dev/null | tail -1 > $O/r02z_bench_prefill2048.json; python -c "
import json; d=json.load(open('$O/r02z_bench_prefill2048.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['next_token'])"
echo "== sanitizer (engine path + prompt attention)"
for tool in memcheck racecheck; do
  LNB_ENGINE=1 LNB_P2P_TIMEOUT_MS=0 timeout 200 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_run.py > $O/r02zz_sanitizer_$tool.txt 2>&1; echo "exit code $?" >> $O/r02zz_sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run\]|exit code" $O/r02zz_sanitizer_$tool.txt | tail -5
done
echo "== launch list prefill"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'sdpa|gemm_tc|rmsnorm|rope|swiglu|resid|gather' -c 150 --csv --log-file $O/r2_launches_prefill.csv python bench.py --config prefill2048 --steps 1 --warmup 1 > $O/r2_prefill_ncu.log 2>&1; python tools/launch_table.py $O/r2_launches_prefill.csv | tee $O/r02z_launches_prefill2048.txt | head -10
} > $O/r02z_final2.log 2>&1
cat $O/r02z_final2.log
