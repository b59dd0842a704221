"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: count, total and share of time."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = r["Kernel Name"]
    rows.append((name, ns))
agg = defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    short = re.sub(r"smr::", "", name)
    short = re.sub(r"\(.*", "", short)
    agg[short][0] += 1
    agg[short][1] += ns
total = sum(v[1] for v in agg.values()) or 1.0
print(f"| kernel | launches | total ms | us/launch | share |")
print(f"|---|---|---|---|---|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ns/1e6:.3f} | {ns/1e3/n:.2f} | {100*ns/total:.1f}% |")
print(f"\ntotal launches {len(rows)}, total kernel time {total/1e6:.3f} ms")
