#!/usr/bin/env bash
# ncu --set full of the round-2 tensor-core kernels: prompt attention + the SwiGLU GEMM (prefill), the FAST batch engine step
set -u
mkdir -p gpurun_out
LNB_P2P_TIMEOUT_MS=0 timeout 500 ncu --set full --clock-control none --import-source on --kernel-name regex:'sdpa_tc_kernel|gemm_tc_kernel' --launch-skip 9 -c 3 -f -o gpurun_out/r02_prefill_kernels python bench.py --config prefill2048 --steps 1 --warmup 1 > gpurun_out/r02_ncu_prefill.log 2>&1
tail -3 gpurun_out/r02_ncu_prefill.log
LNB_P2P_TIMEOUT_MS=0 timeout 500 ncu --set full --clock-control none --import-source on --kernel-name regex:'batch_engine_kernel' --launch-skip 4 -c 1 -f -o gpurun_out/r02_batch_engine_fast python tools/batch_bench.py fast 8 3 > gpurun_out/r02_ncu_batch.log 2>&1
tail -3 gpurun_out/r02_ncu_batch.log
ls -la gpurun_out/*.ncu-rep
