"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv ...`) into
the per-kernel table kept under profiles/ (launches, total / average duration, share of the serialised total).
Usage: python tools/launch_list.py gpurun_out/launches.csv > profiles/rNN_launches_<tag>.txt"""
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = {}
n = 0
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = r[ix["Kernel Name"]].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    n += 1
total = sum(a[1] for a in agg.values())
print(f"# {path}: {n} launches, {total / 1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {cnt:8d} {us:10.1f} {us / cnt:9.2f} {100 * us / total:6.1f}%")
