"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # launches to skip (warm-up)
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((int(r["ID"]), r["Kernel Name"], val * scale))
rows.sort()
rows = rows[skip:]
tot = sum(t for _, _, t in rows)
agg = defaultdict(lambda: [0, 0.0])
for _, name, t in rows:
    short = re.sub(r"<.*", "", name.split("(")[0])
    agg[short][0] += 1
    agg[short][1] += t
print("launches %d   total %.3f ms (cold-cache, serialised: compare SHARES)" % (len(rows), tot / 1e3))
print("%-40s %8s %12s %8s %10s" % ("kernel", "count", "total_us", "share", "avg_us"))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %8d %12.1f %7.1f%% %10.1f" % (name, c, t, 100 * t / tot, t / c))
