#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641"
{
  echo "== decode strict N=$N"; timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --parity-tokens 32 2> gpurun_out/r2f_decode_strict_tp$N.err | tail -1 > gpurun_out/r2f_decode_strict_tp$N.json; echo rc=$?; python -c "
import json
d=json.load(open('gpurun_out/r2f_decode_strict_tp$N.json'))
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d.get('other_acc_mode',{}).get('value'), str(d['parity'])[:300])
"
  echo "== batch8 fast N=$N"; timeout 300 $TR bench.py --gpus $N --config batch8 --acc fast --steps 2 --parity-tokens 4 2> gpurun_out/r2f_batch8_fast_tp$N.err | tail -1 > gpurun_out/r2f_batch8_fast_tp$N.json; echo rc=$?; python -c "
import json
d=json.load(open('gpurun_out/r2f_batch8_fast_tp$N.json'))
print({k:d[k] for k in ('value','n_gpus','decode_ms_per_step','other_acc_mode')})
"
} > gpurun_out/r2f_tp$N.log 2>&1
cat gpurun_out/r2f_tp$N.log
