#!/usr/bin/env python3
"""Generates tests/golden/converged_<config>.npz: block means of CONVERGED renders of BASELINE.json's configurations by the
reference's own CPU engine (oracle/_ref/rz_ref_tool render, built from /root/reference) at the configurations' own
resolution and sample count -- two independent renders A and B each (the engine seeds from the clock), so the fixture
carries its own Monte-Carlo noise floor. Full-resolution float accumulators would be 33 MB each; the tests compare
block means (k x k pixels), so only those are stored (float32) together with the image mean and the spp reached.
Run where /root/reference exists:  python tests/tools/make_golden_images.py [config ...]"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rz_oracle as O  # noqa: E402
from rayzath_b200 import rzs, scenes  # noqa: E402

# name: (world, target spp, max depth, block size, passes per spp estimate)
CASES = {
    "config1_cornell_512": (lambda: scenes.cornell(resolution=(512, 512)), 64, 8, 4),
    "config2_materials_1080p": (lambda: scenes.materials_scene(resolution=(1920, 1080), res=64, cpu_comparable=True), 256, 16, 8),
    "config3_heightfield_1m_1080p": (lambda: scenes.heightfield_scene(resolution=(1920, 1080)), 64, 16, 8),
}


def block_mean(img, k):
    h, w = img.shape[:2]
    img = img[: h // k * k, : w // k * k]
    return img.reshape(h // k, k, w // k, k, img.shape[2]).mean(axis=(1, 3))


def main():
    tmp = tempfile.mkdtemp(prefix="rzb_golden_img_")
    for name in (sys.argv[1:] or list(CASES)):
        make, spp, depth, k = CASES[name]
        w = make()
        path = w.save_reference(os.path.join(tmp, name))
        cam = w.camera_struct()[0]
        W, H = int(cam["width"]), int(cam["height"])
        # passes needed for the target spp: measure the completed paths per pass on a short render first
        probe = os.path.join(tmp, name, "probe.rzs")
        O.ref_tool("render", path, 24, probe, depth, 1, 1, timeout=1800.0)
        acc = np.ascontiguousarray(rzs.read(probe)["accum"]).view(np.float32).reshape(H, W, 4)
        per_pass = float(acc[..., 3].mean()) / 24.0
        passes = int(np.ceil(spp / per_pass)) + depth
        out = {"resolution": np.array([W, H]), "block": np.array([k]), "max_depth": np.array([depth]), "passes": np.array([passes])}
        for tag in ("a", "b"):
            f = os.path.join(tmp, name, "render_%s.rzs" % tag)
            info = O.ref_tool("render", path, passes, f, depth, 1, 1, timeout=3600.0)
            acc = np.ascontiguousarray(rzs.read(f)["accum"]).view(np.float32).reshape(H, W, 4)
            rad = acc[..., :3] / np.maximum(acc[..., 3:4], 1.0)
            out["block_mean_" + tag] = block_mean(rad, k).astype(np.float32)
            out["mean_" + tag] = np.array([rad.mean()], dtype=np.float64)
            out["spp_" + tag] = np.array([acc[..., 3].mean()], dtype=np.float64)
            print(json.dumps({"config": name, "render": tag, "passes": passes, "spp": float(acc[..., 3].mean()),
                              "seconds": info["seconds"], "threads": info.get("threads")}), flush=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "converged_%s.npz" % name), **out)


if __name__ == "__main__":
    main()
