#!/usr/bin/env python3
"""Prints a gpu_variants.py JSON-lines file as a table."""
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(d["workload"][:12], d["variant"].ljust(10), "%.1f" % d["Mrays_s"], "ms %.4f trace %.4f shade %.4f shadow %.4f sort %.4f" % (
        d["ms_per_pass"], d["trace_ms"], d["shade_ms"], d["shadow_ms"], d["sort_ms"]), ["%.5f" % x for x in d["mean_radiance_rel_to_first"]])
