#!/usr/bin/env python3
"""How many closest-hit records differ between the reference trees and the optional SAH trees (all are exact ties)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from rayzath_b200 import capi
import bench
for wl in ("materials_1080p", "heightfield_1m_1080p", "instancing_10m_1080p"):
    res = {}
    for mode in ("reference", "sah"):
        bench.BVH = mode
        w = bench.build_world(wl)
        with capi.Context(0) as c:
            c.set_scene(w.flatten()); c.set_camera(w.camera_struct())
            o, d, nf = c.generate_camera_rays()
            prim = c.trace_closest(o, d, nf)
            if mode == "reference":
                hit = prim["instance"] != capi.NO_INDEX
                rng = np.random.default_rng(5)
                v = rng.normal(size=(int(hit.sum()), 3)).astype(np.float32); v /= np.linalg.norm(v, axis=1, keepdims=True)
                bo = (o[hit] + d[hit] * prim["t"][hit, None] + v * 1e-3).astype(np.float32)
                bnf = np.tile(np.array([[0.0, 3.0e38]], np.float32), (bo.shape[0], 1))
            res[mode] = (prim, c.trace_closest(bo, v, bnf))
    out = {"workload": wl}
    for k, label in ((0, "primary"), (1, "bounce")):
        a, b = res["reference"][k], res["sah"][k]
        diff = (a["instance"] != b["instance"]) | (a["triangle"] != b["triangle"])
        out[label] = {"rays": int(a.shape[0]), "different_record": int(diff.sum()), "of_which_not_a_tie": int((diff & (a["t"] != b["t"])).sum()),
                      "bytes_equal_elsewhere": bool(np.array_equal(a[~diff].view(np.uint8), b[~diff].view(np.uint8)))}
    print(json.dumps(out), flush=True)
