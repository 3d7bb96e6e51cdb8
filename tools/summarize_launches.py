"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of the captured span)."""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = re.sub(r"^void |simx::|\(.*$", "", r["Kernel Name"])
        rows.append((name, ms))
agg = OrderedDict()
for n, ms in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(ms for _, ms in rows)
print(f"launches {len(rows)}, total {tot:.3f} ms")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} n={c:5d}  ms={ms:9.3f}  share={100 * ms / tot:5.1f}%")
