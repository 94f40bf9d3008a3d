#!/usr/bin/env python3
"""Source-line attribution of one kernel of an .ncu-rep: the report's per-instruction `Instructions Executed` joined with the line
table of the matching cubin (nvdisasm -g; build the cubin from the same sources with -lineinfo). The ncu command line offers the
CUDA-source view only in the GUI.  usage: ncu_source_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled kernel substring>"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, kre, cubin, mangled = sys.argv[1:5]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ci, ct, ca = h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("Address")
    body = [r for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
    # the report may list the kernel several times (one block per profiled launch): keep the first
    first = body[0][ca]
    seen, insts = set(), []
    for r in body:
        if r[ca] in seen:
            break
        seen.add(r[ca])
        insts.append((int(r[ca], 16) - int(first, 16), int(r[ci]), float(r[ct] or 0)))
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and mangled in l and ".text." in l)
    line_of, cur, off = {}, None, 0
    for l in dis[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    by_line = collections.Counter()
    lanes = collections.Counter()
    total = 0
    for off, n, thr in insts:
        key = line_of.get(off)
        by_line[key] += n
        lanes[key] += n * thr
        total += n
    print("kernel %s: %d warp instructions, %d SASS instructions, %d with a source line" % (kre, total, len(insts), sum(1 for o, _, _ in insts if line_of.get(o))))
    for key, n in by_line.most_common(40):
        print("%6.2f %%  %4.1f lanes  %s" % (100.0 * n / total, lanes[key] / max(n, 1), "%s:%d" % key if key else "?"))


if __name__ == "__main__":
    main()
