"""Per-region executed-instruction / stall-sample shares from `ncu -i X.ncu-rep --page source --csv > f.csv`.
usage: ncu_src_regions.py f.csv [block]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix['Instructions Executed'] and r[ix['Instructions Executed']].isdigit()]
tot = sum(int(r[ix['Instructions Executed']]) for r in data)
samp = sum(int(r[ix['# Samples']]) for r in data)
print('total warp instructions', tot, 'samples', samp, 'sass lines', len(data))


def op(r):
    t = r[ix['Source']].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]


for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    n = sum(int(r[ix['Instructions Executed']]) for r in seg)
    s = sum(int(r[ix['# Samples']]) for r in seg)
    top = Counter(op(r) for r in seg).most_common(3)
    print(f'{b:5d} {seg[0][ix["Address"]][-5:]} inst {100 * n / tot:5.1f}% samp {100 * s / samp:5.1f}%  {top}')
