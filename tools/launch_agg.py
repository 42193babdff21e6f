"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: count, mean / min / max duration (us), total."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    agg[row["Kernel Name"].split("(")[0][:72]].append(v)
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
    print(f"{k:74s} n={len(v):4d} mean={sum(v) / len(v):9.1f} us  min={min(v):8.1f}  max={max(v):8.1f}  total={sum(v) / 1e3:8.2f} ms")
