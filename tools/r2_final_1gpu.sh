#!/usr/bin/env bash
# round-2 single-GPU evidence with the final binaries -> gpurun_out/r02z_* (copied to profiles/)
set -u
mkdir -p gpurun_out
O=gpurun_out
{
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench decode (default: STRICT)"; timeout 600 python bench.py 2> $O/r02z_bench_1gpu_strict.err | tail -1 > $O/r02z_bench_1gpu_strict.json; python -c "
import json; d=json.load(open('$O/r02z_bench_1gpu_strict.json')); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','roofline','decode_path')}); print(d['e2e']['value'], d['cpu_baseline'], str(d['parity'])[:400])"
echo "== bench batch8 strict"; timeout 600 python bench.py --config batch8 --steps 2 --parity-tokens 16 2>/dev/null | tail -1 > $O/r02z_bench_batch8_strict.json; python -c "
import json; d=json.load(open('$O/r02z_bench_batch8_strict.json')); print(d['value'], d['decode_ms_per_step'], d['other_acc_mode'])"
echo "== bench prefill2048"; timeout 600 python bench.py --config prefill2048 2>/dev/null | tail -1 > $O/r02z_bench_prefill2048.json; python -c "
import json; d=json.load(open('$O/r02z_bench_prefill2048.json')); print(d['value'], d['ms_per_step'], d['roofline'])"
echo "== engine profile"; timeout 400 python tools/engine_prof.py strict,fast > $O/r02_engine_profile.txt 2>&1; tail -16 $O/r02_engine_profile.txt
echo "== ncu launch list of the default bench command"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02z_launches_decode.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02z_ncu_bench.log 2>&1; python tools/launch_table.py $O/r02z_launches_decode.csv > $O/r02z_launches_decode.txt 2>&1; head -14 $O/r02z_launches_decode.txt
echo "== sanitizer"; bash tools/sanitize.sh
echo "== prefill2048 parity vs the oracle (minutes of CPU)"; timeout 560 python bench.py --config prefill2048 --parity --steps 3 2>/dev/null | tail -1 > $O/r02z_bench_prefill2048_parity.json; python -c "
import json; d=json.load(open('$O/r02z_bench_prefill2048_parity.json')); print(d['ms_per_step'], d['parity'])"
} > $O/r02z_final_1gpu.log 2>&1
cat $O/r02z_final_1gpu.log
