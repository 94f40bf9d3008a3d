#!/usr/bin/env python3
"""Find the first pass at which the accumulator of a workload turns non-finite, and where."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rayzath_b200 import capi
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "materials_1080p"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20261018
total = int(sys.argv[3]) if len(sys.argv) > 3 else 320
w = bench.build_world(wl)
with capi.Context(0) as ctx:
    ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct())
    ctx.set_config(1, 1, 16, 0, seed); ctx.reset()
    done = 0
    while done < total:
        ctx.render(8); done += 8
        acc = ctx.read_accum()
        bad = ~np.isfinite(acc).all(axis=2)
        if bad.any():
            ys, xs = np.where(bad)
            print("non-finite after", done, "passes:", len(ys), "pixels; first", [(int(x), int(y), acc[y, x].tolist()) for y, x in zip(ys[:5], xs[:5])])
            break
    else:
        print("finite after", done, "passes")
