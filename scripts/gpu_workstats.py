import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rayzath_b200 import capi
import bench
for wl in ("heightfield_1m_1080p", "materials_1080p"):
    w = bench.build_world(wl)
    with capi.Context(0) as ctx:
        ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct())
        ctx.set_config(1, 1, 16, capi.FLAG_COUNT_WORK, 20261018); ctx.reset()
        ctx.render(64)
        wc = ctx.work_counters()
        print(wl, {k: int(wc[k]) for k in wc.dtype.names})
        seg = int(wc["segments"])
        print("  per segment: nodes %.1f tris %.1f shadow rays %.2f ; per shadow ray: nodes %.1f tris %.1f" % (
            (wc["closest_top_nodes"] + wc["closest_mesh_nodes"]) / seg, wc["closest_triangles"] / seg, wc["shadow_rays"] / seg,
            (wc["shadow_top_nodes"] + wc["shadow_mesh_nodes"]) / max(int(wc["shadow_rays"]), 1), wc["shadow_triangles"] / max(int(wc["shadow_rays"]), 1)))
        print("  lane utilisation of whole-warp batches: closest %.3f shadow %.3f" % (
            wc["closest_lane_work"] / max(int(wc["closest_batch_work"]), 1), wc["shadow_lane_work"] / max(int(wc["shadow_batch_work"]), 1)))
        acc = ctx.read_accum()
        print("  accum finite:", bool(np.isfinite(acc).all()), "spp", float(acc[..., 3].mean()))
