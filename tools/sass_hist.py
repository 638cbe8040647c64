"""SASS evidence that the tensor-core kernels are tcgen05 / TMEM / TMA code: per kernel of libdpmn_b200.so, the count of the
mnemonics /opt/skills/guides/B200_PROFILING.md names (UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU, FFMA2 = packed fp32 FMA).
Usage: python tools/sass_hist.py [dpmn_b200/libdpmn_b200.so] > profiles/rNN_sass_histogram.txt   (no GPU needed)"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dpmn_b200/libdpmn_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UTCATOMSWS", "MUFU", "FFMA2", "HFMA2", "FFMA", "LDG", "STG", "LDS", "STS"]
rows = []
cur, cnt, total = None, None, 0
fi = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur is not None:
            rows.append((cur, cnt, total))
        cur = names[fi].replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("dpmn::", "").replace("<unnamed>::", "")
        fi += 1
        cnt, total = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        total += 1
        op = m.group(2).split(".")[0]
        if op in KEYS:
            cnt[op] += 1
if cur is not None:
    rows.append((cur, cnt, total))
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); kernels with tcgen05 / TMA instructions first")
print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
rows.sort(key=lambda r: (-(r[1]["UTCHMMA"] + r[1]["UTCQMMA"] > 0), -(r[1]["UTMALDG"] > 0), -r[2]))
tot = collections.Counter()
for name, c, n in rows:
    print(f"{name[:64]:64s} {n:6d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    tot.update(c)
print(f"{'TOTAL':64s} {sum(r[2] for r in rows):6d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
