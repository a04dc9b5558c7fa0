"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel table (share of the listed time, launches,
mean / total microseconds, grid, block).  ncu times are serialised and cold-cache: compare SHARES.  Usage: launch_table.py file.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    key = (name, row["Grid Size"], row["Block Size"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    us = v / 1000.0 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1000.0
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values()) or 1.0
print(f"# {sum(a[0] for a in agg.values())} launches, {tot:.1f} us listed (ncu: serialised, cold cache -- compare SHARES)")
print(" share     n    mean_us    total_us  grid block  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * a[1] / tot:6.2f} {a[0]:5d} {a[1] / a[0]:10.2f} {a[1]:11.1f}  {k[1]} {k[2]}  {k[0]}")
