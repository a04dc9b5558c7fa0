"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name
(template arguments kept), launches, total and mean duration, share of the listed time."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""), v * scale))
agg = OrderedDict()
for name, grid, block, us in rows:
    short = re.sub(r"lnb::", "", name)
    key = (short, grid, block)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {len(rows)} launches, {tot:.1f} us listed (ncu: serialised, cold cache -- compare SHARES)")
print(f"{'share':>6} {'n':>5} {'mean_us':>9} {'total_us':>10}  grid block  kernel")
for (name, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * us / tot:6.2f} {n:5d} {us / n:9.2f} {us:10.1f}  {grid} {block}  {name[:150]}")
