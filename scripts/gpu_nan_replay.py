#!/usr/bin/env python3
"""Replay one image row of a workload with the RZB_DEBUG_NAN build (device printf on non-finite values)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rayzath_b200 import capi
import bench
wl, seed, passes, row = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
w = bench.build_world(wl)
with capi.Context(0) as ctx:
    ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct())
    ctx.set_config(1, 1, 16, 0, seed); ctx.set_rows(row, row + 1); ctx.reset()
    ctx.render(passes)
    acc = ctx.read_accum()
    bad = ~np.isfinite(acc).all(axis=2)
    print("non-finite pixels:", [(int(x), int(y)) for y, x in zip(*np.where(bad))])
