#!/usr/bin/env python3
"""Folds `ncu --set full` reports of scripts/profile_target.py into profiles/ncu_summary.json, the per-kernel figures
bench.py reads for `roofline.traffic` (DRAM bytes per launch) and `roofline.issue` (warp instructions per launch, active
threads per instruction): usage  ncu_to_summary.py <workload key> <report.ncu-rep> [<workload key> <report> ...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_summary.json")
FIELDS = {
    "dram_bytes_per_launch": None,  # dram__bytes_read.sum + dram__bytes_write.sum
    "duration_us_under_ncu": "gpu__time_duration.sum",
    "warp_inst_per_launch": "smsp__inst_executed.sum",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_slot_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    summary = json.load(open(OUT)) if os.path.exists(OUT) else {}
    args = sys.argv[1:]
    for key, rep in zip(args[0::2], args[1::2]):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]

        def val(r, name):
            i = hdr.index(name)
            return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)

        entry = summary.setdefault(key, {})
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = "k_trace_paths" if "k_trace_paths" in name else "k_trace_shadow" if "k_trace_shadow" in name else \
                "k_shade" if "k_shade" in name else None
            if short is None or short in seen:
                continue
            seen.add(short)
            entry[short + "_dram_bytes_per_launch"] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            for f, metric in FIELDS.items():
                if metric and metric in hdr:
                    entry[short + "_" + f] = val(r, metric)
        entry["source"] = os.path.basename(rep)
    json.dump(summary, open(OUT, "w"), indent=1)
    print(json.dumps({k: summary[k] for k in args[0::2]}, indent=1))


if __name__ == "__main__":
    main()
