"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for k, us in rows:
    k = re.sub(r"\(.*", "", k)
    c = agg.setdefault(k, [0, 0.0])
    c[0] += 1; c[1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':60s} {'count':>6s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:6d} {us:10.1f} {us/n:8.2f} {100*us/tot:5.1f}%")
