"""Per-instruction stall picture of one kernel from an ncu report (source page, SASS view):
python tools/ncu_source_hot.py report.ncu-rep [top]   -> stall samples by reason, by opcode, and the hottest instructions."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
print("total samples", tot)
agg = Counter()
for r in rows:
    for k in reasons:
        agg[k] += int(r[k] or 0)
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in agg.most_common(10)))
op = Counter(); opx = Counter()
for r in rows:
    o = r["Source"].split()
    name = o[1] if o and o[0].startswith("@") else (o[0] if o else "?")
    name = name.split(".")[0]
    op[name] += int(r["# Samples"] or 0)
    opx[name] += int(r["Instructions Executed"] or 0)
ex = sum(opx.values())
print("by opcode (samples% / executed%):", ", ".join("%s %.1f/%.1f" % (k, 100 * v / tot, 100 * opx[k] / ex) for k, v in op.most_common(14)))
print("hottest instructions:")
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(idx):
    r = rows[i]
    why = sorted(((int(r[k] or 0), k[6:]) for k in reasons), reverse=True)[:3]
    print("%5d %-58s %6s  %s" % (i, r["Source"].strip()[:58], r["# Samples"], " ".join("%s:%d" % (n, v) for v, n in why if v)))
