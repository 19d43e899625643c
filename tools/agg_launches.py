"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = name.replace("void ", "").replace("qlb::", "").replace("GemmCfg", "Cfg")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:9.2f} ms {c:6d} launches {t / c:9.1f} us avg  {100 * t / tot:5.1f}%  {k[:100]}")
print(f"total {tot / 1e3:.2f} ms")
