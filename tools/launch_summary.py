"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    agg[row["Kernel Name"].split("(")[0]].append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:56]:56s} n={len(v):5d} avg={sum(v)/len(v)/1e3:9.2f}us total={sum(v)/1e6:8.2f}ms share={sum(v)/tot:.3f}")
