"""Summarise ncu outputs into small text files for profiles/ (launch shares, key metrics, SASS opcode mix)."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:70]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    out = ["# per-kernel device time from `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised: compare shares)"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:70s} n={v[0]:4d} total={v[1] / 1e6:9.3f} ms  avg={v[1] / v[0] / 1e3:9.1f} us  share={100 * v[1] / tot:5.1f}%")
    return "\n".join(out)


WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_register_spilling', 'sm__cycles_elapsed.max',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']


def raw(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    out = [f"# kernel: {vals[hdr.index('Kernel Name')]}"]
    for i, h in enumerate(hdr):
        if h in WANT or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h):
            out.append(f"{h:85s} {vals[i]} [{units[i]}]")
    return "\n".join(out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        print(launches(sys.argv[2]))
    else:
        print(raw(sys.argv[2]))
