"""Top warp-stall SASS lines of one kernel launch in an .ncu-rep captured with --import-source on.
python tools/ncu_hot.py report.ncu-rep <launch id> [n]"""
import csv
import io
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out))]
hdr = next(r for r in rows if "Source" in r and "Address" in r)
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
num = lambda v: int(v) if v.isdigit() else 0
tot = sum(num(r[i_s]) for r in data)
print(rows[0][1][:100] if rows and len(rows[0]) > 1 else "", "| samples", tot, "| warp instructions", sum(num(r[i_ex]) for r in data))
for r in sorted(data, key=lambda r: -num(r[i_s]))[:n]:
    print(f"{100.0 * num(r[i_s]) / max(tot, 1):5.1f}% {r[i_ex]:>9s}  {r[i_src].strip()[:100]}")
