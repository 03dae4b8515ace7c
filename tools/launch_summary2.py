"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name: total us, launches, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = defaultdict(lambda: [0.0, 0])
for n, us in rows:
    agg[n][0] += us
    agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot / 1e3:.3f} ms over {len(rows)} launches")
for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}% {c:5d}x  {n[:110]}")
