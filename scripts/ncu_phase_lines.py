#!/usr/bin/env python3
"""Like ncu_source_lines.py, but every SASS instruction is attributed to the line of the TRAVERSAL code it was inlined into
(nvdisasm -gi inline chains: the outermost rzb_traverse.cuh line below `--below`, i.e. inside trav_begin / trav_round, else the
kernel's own line), so that helper lines (fmul, fadd ...) are charged to their call sites.
usage: ncu_phase_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled kernel substring> [--below LINE] [--top N]"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, kre, cubin, mangled = sys.argv[1:5]
    below = int(sys.argv[sys.argv.index("--below") + 1]) if "--below" in sys.argv else 10 ** 9
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 60
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ci, ct, ca = h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("Address")
    body = [r for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
    first = body[0][ca]
    seen, insts = set(), []
    for r in body:
        if r[ca] in seen:
            break
        seen.add(r[ca])
        insts.append((int(r[ca], 16) - int(first, 16), int(r[ci]), float(r[ct] or 0)))
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and mangled in l and ".text." in l)
    key_of, chain, fresh = {}, [], True
    for l in dis[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            fresh = True
            trav = [c for c in chain if c[0] == "rzb_traverse.cuh" and c[1] < below]
            key = trav[-1] if trav else (chain[-1] if chain else None)
            key_of[int(m.group(1), 16)] = key
    by, lanes, total = collections.Counter(), collections.Counter(), 0
    for off, n, thr in insts:
        k = key_of.get(off)
        by[k] += n
        lanes[k] += n * thr
        total += n
    print("kernel %s: %d warp instructions, %d SASS instructions" % (kre, total, len(insts)))
    for k, n in by.most_common(top):
        print("%6.2f %%  %4.1f lanes  %s" % (100.0 * n / total, lanes[k] / max(n, 1), "%s:%d" % k if k else "?"))


if __name__ == "__main__":
    main()
