#!/usr/bin/env python3
"""Where a kernel's warps wait: the warp-stall samples of an .ncu-rep (ncu --set full --import-source on) per source line
(nvdisasm -gi inline chains; the outermost line inside `--file`, default the kernel's own file) with the dominant stall reasons.
usage: ncu_stall_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled kernel substring> [--file NAME] [--top N]"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, kre, cubin, mangled = sys.argv[1:5]
    fname = sys.argv[sys.argv.index("--file") + 1] if "--file" in sys.argv else "rzb_kernels.cuh"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]
    ca, cs, ci = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_")]
    body = [r for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
    first = body[0][ca]
    seen, insts = set(), []
    for r in body:
        if r[ca] in seen:
            break
        seen.add(r[ca])
        insts.append((int(r[ca], 16) - int(first, 16), int(r[cs] or 0), int(r[ci]), {c: int(r[i] or 0) for i, c in stall_cols}, r[h.index("Source")]))
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
            own = [c for c in chain if c[0] == fname]
            key_of[int(m.group(1), 16)] = own[-1] if own else (chain[-1] if chain else None)
    samples, reasons, execs = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
    total = 0
    for off, n, ex, st, _ in insts:
        k = key_of.get(off)
        samples[k] += n
        execs[k] += ex
        total += n
        for c, v in st.items():
            reasons[k][c] += v
    print("kernel %s: %d stall samples, %d warp instructions" % (kre, total, sum(execs.values())))
    for k, n in samples.most_common(top):
        rs = ", ".join("%s %d%%" % (c[6:], 100 * v // max(n, 1)) for c, v in reasons[k].most_common(3))
        print("%6.2f %% samples  %6.2f %% instr  %-26s %s" % (100.0 * n / max(total, 1), 100.0 * execs[k] / max(sum(execs.values()), 1),
                                                          "%s:%d" % k if k else "?", rs))


if __name__ == "__main__":
    main()
