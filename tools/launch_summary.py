"""Summarise an ncu launch list (gpu__time_duration.sum csv): python tools/launch_summary.py launches.csv"""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = 0.0
agg = {}
for x in csv.DictReader(lines):
    name = x["Kernel Name"][:100]
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
for k, a in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{a[1]:10.1f} us {100 * a[1] / tot:5.1f}% x{a[0]:3d} {k}")
print(f"{tot:10.1f} us total")
