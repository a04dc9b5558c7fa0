#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'sdpa|gemm_tc|rmsnorm|rope|swiglu|resid|gather' -c 150 --csv --log-file gpurun_out/r2_launches_prefill.csv python bench.py --config prefill2048 --steps 1 --warmup 1 > gpurun_out/r2_prefill_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows = []
with open("gpurun_out/r2_launches_prefill.csv") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]); key = (name, row["Grid Size"], row["Block Size"])
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    us = v / 1000.0 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1000.0
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {sum(a[0] for a in agg.values())} launches, {tot:.1f} us listed")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*a[1]/tot:6.2f} {a[0]:5d} {a[1]/a[0]:10.2f} {a[1]:10.1f}  {k[1]} {k[2]}  {k[0]}")
PY
