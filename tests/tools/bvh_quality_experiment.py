#!/usr/bin/env python3
"""CPU experiment (oracle work counters): nodes visited / triangles tested per ray on the reference's tree versus
the optional SAH tree, for primary rays and for diffuse bounce rays leaving the primary hit points."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import rz_oracle as O
from rayzath_b200 import scenes

def rays_for(w, flat, n=200_000, seed=3):
    cam = w.camera_struct()
    o, d, nf = O.camera_rays(cam)
    rng = np.random.default_rng(seed)
    pick = rng.choice(o.shape[0], size=min(n, o.shape[0]), replace=False)
    o, d, nf = o[pick], d[pick], nf[pick]
    hits = O.trace_closest(O.Scene(flat), o, d, nf)
    ok = hits["instance"] != 0xFFFFFFFF
    p = o[ok] + d[ok] * hits["t"][ok, None]
    # uniform directions on the sphere, pushed off the surface a little along the direction
    v = rng.normal(size=p.shape).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    bo = (p + v * 1e-3).astype(np.float32)
    bnf = np.tile(np.array([[0.0, 3.4e38]], np.float32), (bo.shape[0], 1))
    return (o, d, nf), (bo, v.astype(np.float32), bnf)

def stats(flat, rays):
    t = time.time()
    hits, st = O.trace_closest(O.Scene(flat), *rays, stats=True)
    n = rays[0].shape[0]
    return hits, {"nodes": float(st["top_nodes"] + st["mesh_nodes"]) / n, "tris": float(st["triangles"]) / n, "cpu_s": round(time.time() - t, 2)}

for name, make in (("materials", lambda: scenes.materials_scene(resolution=(640, 360), res=64)),
                   ("heightfield_1m", lambda: scenes.heightfield_scene(resolution=(640, 360)))):
    w = make()
    flat_ref = w.flatten()
    prim, bounce = rays_for(w, flat_ref)
    res = {}
    for label, builder in (("reference", "reference"), ("sah8", ("sah", 8)), ("sah4", ("sah", 4)), ("sah2", ("sah", 2))):
        for m in w.meshes:
            m.bvh_builder = builder
        t = time.time(); flat = w.flatten(); bt = time.time() - t
        hp, sp = stats(flat, prim)
        hb, sb = stats(flat, bounce)
        res[label] = (hp, hb)
        same_p = np.mean((hp["instance"] == res["reference"][0]["instance"]) & (hp["triangle"] == res["reference"][0]["triangle"]) & (hp["t"] == res["reference"][0]["t"]))
        same_b = np.mean((hb["instance"] == res["reference"][1]["instance"]) & (hb["triangle"] == res["reference"][1]["triangle"]) & (hb["t"] == res["reference"][1]["t"]))
        print(json.dumps({"scene": name, "tree": label, "nodes": int(flat["mesh_nodes"].shape[0]), "flatten_s": round(bt, 2),
                          "primary": sp, "bounce": sb, "same_hit_primary": float(same_p), "same_hit_bounce": float(same_b)}), flush=True)
