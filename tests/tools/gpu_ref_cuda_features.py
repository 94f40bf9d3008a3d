#!/usr/bin/env python3
"""Feature-by-feature comparison of the tone-mapped image of the reference's CUDA engine with this repo's
(CUDA-engine semantics): mean RGB8 per variant of the materials scene."""
import json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rayzath_b200 import capi, rzs, scenes
TOOL = os.path.join(ROOT, "oracle", "_ref", "rz_ref_tool_cuda")

def variant(keep, lights):
    w = scenes.materials_scene(resolution=(640, 360), res=32)
    w.instances = [i for i in w.instances if i.name in keep]
    for k, i in enumerate(w.instances): i.index = k
    if "sun" not in lights: w.direct_lights = []
    if "spots" not in lights: w.spot_lights = []
    return w

VARIANTS = {
    "ground_sky_only": (["ground"], []),
    "ground_sun": (["ground"], ["sun"]),
    "ground_spots": (["ground"], ["spots"]),
    "ground_mirror_sun": (["ground", "mirror ball"], ["sun"]),
    "ground_glossy_sun": (["ground", "glossy torus"], ["sun"]),
    "ground_glass_sun": (["ground", "glass cylinder"], ["sun"]),
    "ground_fog_sun": (["ground", "fog ball"], ["sun"]),
    "ground_gold_sun": (["ground", "gold ball"], ["sun"]),
    "fog_sun": (["fog ball"], ["sun"]),
    "fog_sky_only": (["fog ball"], []),
    "ground_fog_sky_only": (["ground", "fog ball"], []),
    "all": (["ground", "mirror ball", "glossy torus", "glass cylinder", "fog ball", "gold ball"], ["sun", "spots"]),
}
tmp = tempfile.mkdtemp(prefix="rzb_feat_")
# RZ_REF_CUDA_VARIANT=_nofma compares against the copy built with `make ref_cuda VARIANT=_nofma NVEXTRA=-fmad=false`
TOOL += os.environ.get("RZ_REF_CUDA_VARIANT", "")
SAVE = os.environ.get("RZ_SAVE_IMAGES")
for name, (keep, lights) in VARIANTS.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]: continue
    w = variant(keep, lights)
    path = w.save_reference(os.path.join(tmp, name))
    # the reference draws its 256 seeds once per renderWorld call (cuda_kernel_data.cu:10-18), so the passes of one
    # call reuse the same random numbers per (pixel, depth): few calls x many passes = few independent samples
    calls, rpp = int(os.environ.get("RZ_CALLS", 9)), int(os.environ.get("RZ_RPP", 64))
    out = os.path.join(tmp, name, "cuda.rzs")
    r = subprocess.run([TOOL, "rendercuda", path, str(calls), str(rpp), out, "16", "1", "1", "1"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print(json.dumps({"variant": name, "error": r.stderr[-300:]})); continue
    ref = rzs.read(out); W, H = int(ref["resolution"][0]), int(ref["resolution"][1])
    a = ref["rgba8"].reshape(H, W, 4)[..., :3].astype(np.float64)
    with capi.Context(0) as ctx:
        ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct()); ctx.set_config(1, 1, 16, 0, 5); ctx.reset()
        ctx.render(calls * rpp)
        b = ctx.resolve()[0][..., :3].astype(np.float64)
    if SAVE:
        os.makedirs(SAVE, exist_ok=True)
        np.savez_compressed(os.path.join(SAVE, name + os.environ.get("RZ_REF_CUDA_VARIANT", "") + ".npz"), ref=a.astype(np.uint8), ours=b.astype(np.uint8))
    # per-row-band means (top = sky, bottom = ground) to localise differences
    bands = [(0, H // 3), (H // 3, 2 * H // 3), (2 * H // 3, H)]
    print(json.dumps({"variant": name, "calls": calls, "rpp": rpp, "mean_ref": np.round(a.mean(axis=(0, 1)), 2).tolist(), "mean_ours": np.round(b.mean(axis=(0, 1)), 2).tolist(),
                      "ratio": np.round(b.mean(axis=(0, 1)) / np.maximum(a.mean(axis=(0, 1)), 1e-9), 3).tolist(),
                      "band_ratio": [round(float(b[y0:y1].mean() / max(a[y0:y1].mean(), 1e-9)), 3) for y0, y1 in bands]}), flush=True)
