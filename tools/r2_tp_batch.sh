#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
{
  echo "== tp_check tiny N=$N"; timeout 400 $TR tools/tp_check.py tiny 2>&1 | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ok', d['ok'])
for k in ('strict','fast','strict_p2p','fast_p2p'):
    print(k, d[k]['bit_exact_vs_oracle_tp_order'], d[k]['max_abs_vs_oracle_tp_order'], 'batch3', d[k]['batch3'])
"
  echo "== batch8 bench FAST N=$N"; timeout 600 $TR bench.py --gpus $N --config batch8 --acc fast --steps 2 --parity-tokens 8 > gpurun_out/r2_bench_batch8_fast_tp$N.json 2> gpurun_out/r2_bench_batch8_fast_tp$N.err; echo rc=$?; tail -c 1500 gpurun_out/r2_bench_batch8_fast_tp$N.json; tail -3 gpurun_out/r2_bench_batch8_fast_tp$N.err
  echo "== batch8 bench N=$N"; timeout 600 $TR bench.py --gpus $N --config batch8 --steps 2 --parity-tokens 8 > gpurun_out/r2_bench_batch8_tp$N.json 2> gpurun_out/r2_bench_batch8_tp$N.err; echo rc=$?; tail -c 1800 gpurun_out/r2_bench_batch8_tp$N.json; tail -3 gpurun_out/r2_bench_batch8_tp$N.err
} > gpurun_out/r2_tp_batch$N.log 2>&1
cat gpurun_out/r2_tp_batch$N.log
