#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'engine|gemv|sdpa|rms_scale|gather|set_state|argmax|p2p' -c 700 --csv --log-file gpurun_out/r02z_launches_decode.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02z_ncu_bench.log 2>&1
python tools/launch_table.py gpurun_out/r02z_launches_decode.csv | tee gpurun_out/r02z_launches_decode.txt | head -20
