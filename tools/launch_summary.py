"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: python tools/launch_summary.py in.csv [out.txt] [header text]"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], ms))
tot = sum(ms for _, ms in rows)
agg = collections.OrderedDict()
for k, ms in rows:
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + ms)
out = [sys.argv[3]] if len(sys.argv) > 3 else []
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    out.append("%-80s n=%4d %12.3f ms %6.1f%%" % (k[:80], n, t, 100 * t / tot))
out.append("total %d launches, %.3f ms" % (len(rows), tot))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
