#!/usr/bin/env bash
# tensor-parallel check + bench at N GPUs (N = number of visible GPUs).   gpurun --gpus N --timeout 1500 -- 'bash tools/r2_tp.sh'
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r2_tp${N}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
{
  echo "== tp_check tiny, N=$N (engine default)"; timeout 400 $TR tools/tp_check.py tiny 2>&1 | tail -3
  echo "== bench N=$N engine"; timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 --parity-tokens ${PAR:-24} > gpurun_out/r2_bench_tp${N}_engine.json 2> gpurun_out/r2_bench_tp${N}_engine.err; echo rc=$?; tail -c 2500 gpurun_out/r2_bench_tp${N}_engine.json; tail -3 gpurun_out/r2_bench_tp${N}_engine.err
  echo "== bench N=$N chain (LNB_ENGINE=0)"; LNB_ENGINE=0 timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 --no-cpu > gpurun_out/r2_bench_tp${N}_chain.json 2> gpurun_out/r2_bench_tp${N}_chain.err; echo rc=$?; head -c 700 gpurun_out/r2_bench_tp${N}_chain.json; tail -3 gpurun_out/r2_bench_tp${N}_chain.err
} > "$OUT" 2>&1
tail -40 "$OUT"
