#!/usr/bin/env bash
# round-2 multi-GPU evidence: tensor-parallel parity check, config 5 (batch8) in both modes, the default decode line
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
{
  echo "== tp_check tiny N=$N"; timeout 400 $TR tools/tp_check.py tiny 2>&1 | grep "^{" | tee gpurun_out/r2f_tp_check_$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('ok', d['ok'])
for k in ('strict','fast','strict_p2p','fast_p2p'):
    print(k, d[k]['bit_exact_vs_oracle_tp_order'], d[k]['max_abs_vs_oracle_tp_order'], 'batch3', d[k]['batch3'])
"
  for acc in strict fast; do
    echo "== batch8 $acc N=$N"; timeout 600 $TR bench.py --gpus $N --config batch8 --acc $acc --steps 2 --parity-tokens 8 2> gpurun_out/r2f_batch8_${acc}_tp$N.err | tail -1 > gpurun_out/r2f_batch8_${acc}_tp$N.json; echo rc=$?; python -c "
import json,sys
d=json.load(open('gpurun_out/r2f_batch8_${acc}_tp$N.json'))
print({k:d[k] for k in ('value','n_gpus','decode_ms_per_step','e2e','other_acc_mode')}, d['config']['parallelism'], d['config']['collective'], [p['oracle_tokens_equal'] for p in d['parity']['per_sequence']])
"
  done
  echo "== decode strict N=$N"; timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/r2f_decode_strict_tp$N.err | tail -1 > gpurun_out/r2f_decode_strict_tp$N.json; echo rc=$?; python -c "
import json
d=json.load(open('gpurun_out/r2f_decode_strict_tp$N.json'))
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e','decode_path','parity')})
"
} > gpurun_out/r2f_tp$N.log 2>&1
cat gpurun_out/r2f_tp$N.log
