#!/usr/bin/env python3
"""Bound for the exact-decision closest-hit kernel: the conservative (FAST) kernels on the REFERENCE trees."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rayzath_b200 import capi
import bench
for wl in ("materials_1080p", "heightfield_1m_1080p"):
    w = bench.build_world(wl)
    flat = w.flatten()
    for label, flags in (("exact", 0), ("conservative", capi.SCENE_OWN_TREES)):
        f2 = dict(flat); f2["scene_flags"] = np.array([flags], np.uint32)
        with capi.Context(0) as ctx:
            ctx.set_scene(f2); ctx.set_camera(w.camera_struct()); ctx.set_config(1, 1, 16, 0, 20261018); ctx.reset()
            ctx.render(64); ctx.render(256)
            st = ctx.render_stats()
            print(json.dumps({"workload": wl, "boxes": label, "trace_ms": round(float(st["last_trace_ms"]), 4),
                              "shade_ms": round(float(st["last_shade_ms"]), 4), "shadow_ms": round(float(st["last_shadow_ms"]), 4)}), flush=True)
