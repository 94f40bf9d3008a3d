#!/usr/bin/env python3
"""Short program for ncu: set up a workload and render a few passes (no timing claims are taken from runs under ncu)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayzath_b200 import capi
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="heightfield_1m_1080p")
ap.add_argument("--passes", type=int, default=12)
ap.add_argument("--bvh", default="reference", choices=["reference", "sah"])
ap.add_argument("--primary", type=int, default=0, help="also run the ray-set entry point on the primary rays N times")
a = ap.parse_args()
bench.BVH = a.bvh
w = bench.build_world(a.workload)
with capi.Context(0) as ctx:
    ctx.set_scene(w.flatten())
    ctx.set_camera(w.camera_struct())
    ctx.set_config(1, 1, bench.MAX_DEPTH, 0, 20261018)
    ctx.reset()
    ctx.render(a.passes)
    ctx.synchronize()
    if a.primary:
        o, d, nf = ctx.generate_camera_rays()
        for _ in range(a.primary):
            ctx.trace_closest(o, d, nf)
    print("ok", ctx.render_stats())
