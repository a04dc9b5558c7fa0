#!/usr/bin/env bash
# Times the unmodified Go reference on the synthetic checkpoint.  See README.md.
#   usage: baseline/go/run_reference.sh <reference checkout> <scratch model dir> [tiny]
set -euo pipefail
REF=${1:?path to a checkout of adalkiran/llama-nuts-and-bolts}
DIR=${2:?scratch directory for the synthetic model (16 GB for the 8B architecture)}
SIZE=${3:-}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
command -v go >/dev/null || { echo "no Go toolchain on PATH: skipping (bench.py --impl reference times the C restatement instead)"; exit 0; }
make -C "$ROOT/llama-nuts-and-bolts_b200/csrc" -s
make -C "$ROOT/host" -s
mkdir -p "$DIR"
"$ROOT/host/lnb_generate" --write-synthetic "$DIR" $SIZE
mkdir -p "$REF/cmd/lnb_bench"
cp "$HERE/lnb_bench_main.go" "$REF/cmd/lnb_bench/main.go"
(cd "$REF" && go run ./cmd/lnb_bench "$DIR" 136)
