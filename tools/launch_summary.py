"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of step."""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(r[1] for r in rows)
agg = defaultdict(lambda: [0.0, 0])
for n, ns, g, b in rows:
    agg[n][0] += ns
    agg[n][1] += 1
print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
for n, (ns, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n[:70]:70s} {c:8d} {ns / 1e3:10.1f} {ns / 1e3 / c:9.2f} {100 * ns / tot:6.1f}%")
if "--detail" in sys.argv:
    for i, (n, ns, g, b) in enumerate(rows):
        print(i, n[:60], f"{ns / 1e3:.1f}us", g, b)
