"""Chunked view of an `ncu --page source --csv` dump: per run of CH SASS instructions, the share of stall samples,
executed warp-instructions, samples per executed instruction, dominant opcodes and stall reasons (dev tool)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('instructions', len(data), 'samples', tot)
for c0 in range(0, len(data), CH):
    ch = data[c0:c0 + CH]
    s = sum(int(r[ix['# Samples']]) for r in ch)
    ex = sum(int(r[ix['Instructions Executed']]) for r in ch)
    ops = collections.Counter()
    for r in ch:
        t = r[ix['Source']].split()
        op = (t[1] if t and t[0].startswith('@') else t[0]).split('.')[0] if t else '?'
        ops[op] += 1
    st = collections.Counter()
    for r in ch:
        for h in stall_cols:
            st[h[6:]] += int(r[ix[h]] or 0)
    top = ' '.join(f'{k}:{v}' for k, v in st.most_common(4))
    print(f'{c0:5d} samp {100 * s / tot:5.2f}% exec {ex / 1e6:7.2f}M  spi {s / max(ex, 1) * 1e6:7.1f} | '
          f'{" ".join(f"{k}{v}" for k, v in ops.most_common(5)):50s} | {top}')
