#!/usr/bin/env python3
"""Run the REFERENCE'S OWN CUDA engine (oracle/_ref/rz_ref_tool_cuda: cuda_*.cu compiled in place for sm_100a) on
the bench workloads: the GPU comparand. Prints one JSON object per workload with its Mrays/s and a statistical
comparison of its tone-mapped image with this repo's (CUDA-engine semantics, same number of passes)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rayzath_b200 import capi, rzs  # noqa: E402
import bench  # noqa: E402

TOOL = os.path.join(ROOT, "oracle", "_ref", "rz_ref_tool_cuda")


def main():
    tmp = tempfile.mkdtemp(prefix="rzb_refcuda_")
    workloads = sys.argv[1:] or ["materials_1080p", "heightfield_1m_1080p"]
    for wl in workloads:
        w = bench.build_world(wl)
        path = w.save_reference(os.path.join(tmp, wl))
        calls, rpp = 9, 32
        out = os.path.join(tmp, wl, "cuda.rzs")
        r = subprocess.run([TOOL, "rendercuda", path, str(calls), str(rpp), out, str(bench.MAX_DEPTH), "1", "1", "1"],
                           capture_output=True, text=True, timeout=1200)
        if r.returncode != 0:
            print(json.dumps({"workload": wl, "error": r.stderr[-600:]}), flush=True)
            continue
        info = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        ref = rzs.read(out)
        W, H = int(ref["resolution"][0]), int(ref["resolution"][1])
        ref_rgba = ref["rgba8"].reshape(H, W, 4)
        passes = calls * rpp
        with capi.Context(0) as ctx:
            ctx.set_scene(w.flatten())
            ctx.set_camera(w.camera_struct())
            ctx.set_config(1, 1, bench.MAX_DEPTH, capi.FLAG_NONE, 99)
            ctx.reset()
            ctx.render(passes)
            ctx.synchronize()
            st = ctx.render_stats()
            rgba, depth, _ = ctx.resolve(want_depth=True)
        a, b = ref_rgba[..., :3].astype(np.float64), rgba[..., :3].astype(np.float64)
        depth_ref = ref["depth"].reshape(H, W)
        finite = np.isfinite(depth_ref) & np.isfinite(depth)
        print(json.dumps({
            "workload": wl, "reference_cuda_mrays_s": info["timed_rays"] / info["seconds"] / 1e6,
            "reference_cuda_ms_per_pass": info["seconds"] / max(info["timed_calls"] * rpp, 1) * 1e3,
            "ours_ms_per_pass (device)": float(st["last_render_ms"]) / passes,
            "ours_mrays_s": passes * W * H / (float(st["last_render_ms"]) * 1e-3) / 1e6,
            "passes": passes, "mean_rgb8_reference_cuda": a.mean(axis=(0, 1)).tolist(), "mean_rgb8_ours": b.mean(axis=(0, 1)).tolist(),
            "rmse_rgb8": float(np.sqrt(((a - b) ** 2).mean())),
            "first_pass_depth_rel_diff_median": float(np.median(np.abs(depth_ref[finite] - depth[finite]) / np.maximum(np.abs(depth_ref[finite]), 1e-6))),
        }), flush=True)


if __name__ == "__main__":
    main()
