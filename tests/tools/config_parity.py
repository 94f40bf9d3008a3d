#!/usr/bin/env python3
"""Converged-image parity of BASELINE.json's configurations at their real resolutions: the B200 path (CPU
semantics flag) against the reference's own CPU engine (oracle/_ref/rz_ref_tool render) at EQUAL passes, with the
Monte-Carlo noise floor taken from two independent reference renders. Writes one JSON object per configuration.
Run on the GPU box:  python tests/tools/config_parity.py > gpurun_out/config_parity.jsonl"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rz_oracle as O  # noqa: E402
from rayzath_b200 import capi, rzs, scenes  # noqa: E402

CASES = [
    # name, world, passes, max_depth, note
    ("config1_cornell_512", lambda: scenes.cornell(resolution=(512, 512)), 512, 8,
     "Cornell box 512x512, depth 8, 512 passes (>= 64 spp)"),
    ("config2_materials_1080p", lambda: scenes.materials_scene(resolution=(1920, 1080), res=64, cpu_comparable=True), 96, 16,
     "materials + lights 1920x1080, NEE 1+1, 96 passes (the reference CPU engine bounds the spp; scattering medium off: "
     "the CPU engine has none)"),
    ("config3_heightfield_1m_1080p", lambda: scenes.heightfield_scene(resolution=(1920, 1080)), 64, 16,
     "1M-triangle height field with texture / normal / roughness maps and DoF, 1920x1080, 64 passes"),
]


def radiance(acc):
    return acc[..., :3] / np.maximum(acc[..., 3:4], 1.0)


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))


def block_mean(img, k):
    h, w = img.shape[:2]
    img = img[: h // k * k, : w // k * k]
    return img.reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))


def main():
    tmp = tempfile.mkdtemp(prefix="rzb_cfg_")
    only = sys.argv[1:]
    for name, make, passes, depth, note in CASES:
        if only and name not in only:
            continue
        w = make()
        path = w.save_reference(os.path.join(tmp, name))
        cam = w.camera_struct()[0]
        W, H = int(cam["width"]), int(cam["height"])
        refs, ref_secs = [], []
        for tag in ("a", "b"):
            out = os.path.join(tmp, name, "render_%s.rzs" % tag)
            info = O.ref_tool("render", path, passes, out, depth, 1, 1, timeout=1200.0)
            refs.append(np.ascontiguousarray(rzs.read(out)["accum"]).view(np.float32).reshape(H, W, 4))
            ref_secs.append(info["seconds"])
        t0 = time.time()
        with capi.Context(0) as ctx:
            ctx.set_scene(w.flatten())
            ctx.set_camera(w.camera_struct())
            ctx.set_config(1, 1, depth, capi.FLAG_CPU_SEMANTICS, 2024)
            ctx.reset()
            ctx.render(passes)
            acc = ctx.read_accum()
            gpu_ms = float(ctx.render_stats()["last_render_ms"])
        A, B, G = radiance(refs[0]), radiance(refs[1]), radiance(acc)
        sigma = rel_rmse(A, B)
        bs = rel_rmse(block_mean(A, 8), block_mean(B, 8))
        res = {
            "config": name, "note": note, "resolution": [W, H], "passes": passes, "max_depth": depth,
            "spp_reference": float(refs[0][..., 3].mean()), "spp_gpu": float(acc[..., 3].mean()),
            "rel_rmse_gpu_vs_ref": rel_rmse(G, A), "rel_rmse_ref_vs_ref (noise floor sigma)": sigma,
            "block8_rel_rmse_gpu_vs_refmean": rel_rmse(block_mean(G, 8), block_mean(0.5 * (A + B), 8)),
            "block8_rel_rmse_ref_vs_ref": bs,
            "mean_radiance_gpu": float(G.mean()), "mean_radiance_ref": float(0.5 * (A.mean() + B.mean())),
            "tolerance": "rel_rmse <= 1.25 sigma + 0.01; block8 <= 1.25 block sigma + 0.02; mean within 3 %",
            "within_tolerance": bool(rel_rmse(G, A) <= 1.25 * sigma + 0.01 and
                                     rel_rmse(block_mean(G, 8), block_mean(0.5 * (A + B), 8)) <= 1.25 * bs + 0.02 and
                                     abs(G.mean() - 0.5 * (A.mean() + B.mean())) / A.mean() < 0.03),
            "gpu_render_ms": gpu_ms, "reference_cpu_seconds": ref_secs, "reference_threads": info.get("threads"),
        }
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
