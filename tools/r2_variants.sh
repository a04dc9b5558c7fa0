#!/usr/bin/env bash
# A/B of engine build variants (build_variants/liblnb_X.so, built locally): kbench fast[,strict] with each.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2_variants.log
cp llama-nuts-and-bolts_b200/liblnb.so /tmp/liblnb_orig.so
{
  for v in ${VARIANTS:-A B C D}; do
    cp build_variants/liblnb_$v.so llama-nuts-and-bolts_b200/liblnb.so
    echo "== variant $v"; LNB_ENGINE_PF_KB=${PF:-0} timeout 200 python tools/kbench.py ${MODES:-fast} 2>&1 | tail -${TAIL:-2}
  done
} > "$OUT" 2>&1
cp /tmp/liblnb_orig.so llama-nuts-and-bolts_b200/liblnb.so
tail -40 "$OUT"
