#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) into the text summary kept under profiles/: headline metrics per captured launch,
instruction mix and the hottest SASS blocks with their average active threads. Usage: summarize_ncu.py rep [rep...]"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']


def main():
    for rep in sys.argv[1:]:
        print("=" * 100)
        print("report:", rep)
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("--- launch:", r[hdr.index('Kernel Name')][:90])
            for w in WANT:
                if w in hdr:
                    print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
            for i, h in enumerate(hdr):
                if 'issue_stalled' in h and h.endswith('per_warp_active.pct'):
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v > 4:
                        print("  stall %-66s %.1f %%" % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
        if not hi:
            continue
        h = rows[hi[0]]
        body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
        ci, cs, ct = h.index('Instructions Executed'), h.index('Source'), h.index('Avg. Threads Executed')
        tot = sum(int(r[ci]) for r in body if len(r) > ci and r[ci].isdigit())
        byop = collections.Counter()
        blocks, cur = [], None
        for k, r in enumerate(body):
            if len(r) <= ci or not r[ci].isdigit():
                continue
            c = int(r[ci])
            t = r[cs].strip().split()
            op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            byop[op] += c
            if cur and cur['c'] == c:
                cur['n'] += 1
                cur['ops'].append(op)
            else:
                cur = {'c': c, 'n': 1, 'start': k, 'thr': r[ct], 'ops': [op]}
                blocks.append(cur)
        print("  warp instructions (first launch): %d" % tot)
        print("  mix: " + ' '.join('%s %.1f%%' % (op, 100 * c / tot) for op, c in byop.most_common(14)))
        print("  hottest SASS blocks (consecutive instructions with equal execution count):")
        for b in blocks:
            if b['c'] * b['n'] > 0.02 * tot:
                oc = collections.Counter(b['ops'])
                print("    sass#%-5d n=%-3d executed=%-10d share=%5.1f%% avg_active_threads=%-4s %s" % (
                    b['start'], b['n'], b['c'], 100 * b['c'] * b['n'] / tot, b['thr'], dict(oc.most_common(5))))


if __name__ == "__main__":
    main()
