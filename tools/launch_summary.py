"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{'us':>10s} {'share':>6s} {'n':>5s} {'avg us':>9s}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:5d} {t / n:9.1f}  {k[:120]}")
    print(f"{tot:10.1f} total")


if __name__ == "__main__":
    main(sys.argv[1])
