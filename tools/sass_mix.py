"""Aggregate an `ncu --page source --csv` dump by SASS opcode: executed warp-instructions and stall samples."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ei, ss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
agg = collections.Counter(); smp = collections.Counter(); tot = 0; stot = 0
for r in rows[2:]:
    try:
        n = int(r[ei]); s = int(r[ss])
    except Exception:
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    op = op.rstrip(';')
    key = op.split('.')[0]
    if key in ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL'):
        key = '.'.join(op.split('.')[:1] + [x for x in op.split('.')[1:] if x in ('128', '64', 'U8', 'U16')])
    agg[key] += n; smp[key] += s; tot += n; stot += s
de = float(sys.argv[2]) if len(sys.argv) > 2 else None
for op, n in agg.most_common(40):
    extra = f" {n * 32 / de:7.2f}/DE" if de else ""
    print(f"{op:14s} {n:12d} {100 * n / tot:5.1f}%  samples {100 * smp[op] / max(stot, 1):5.1f}%{extra}")
print("total", tot, (f"{tot * 32 / de:.1f}/DE" if de else ""))
