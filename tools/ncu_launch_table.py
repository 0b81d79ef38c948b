"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (dev tool).
usage: python tools/ncu_launch_table.py launches.csv [--by-grid]"""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ki, vi, ui, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
by_grid = "--by-grid" in sys.argv
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    name = r[ki].split("(")[0]
    d[(name, r[gi]) if by_grid else (name, "")].append(v)
tot = sum(sum(v) for v in d.values())
print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised: compare shares)")
for (k, g), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:44s} {g:>16s} n={len(v):4d} total={sum(v) / 1e3:9.3f} ms  avg={sum(v) / len(v):9.1f} us  share={sum(v) / tot * 100:5.1f}%")
