#!/usr/bin/env python3
"""Developer probe (GPU): where an e2e frame of bench.py spends its time (host wall clock, synchronised after each call)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from rayzath_b200 import capi

w = bench.build_world(sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD)
flat, cam = w.flatten(), w.camera_struct()
H, W = int(cam[0]["height"]), int(cam[0]["width"])
host = {k: np.ascontiguousarray(v) for k, v in flat.items()}
bench.pin(list(host.values()))
rgba = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()
depth = torch.empty((H, W), dtype=torch.float32, pin_memory=True).numpy()
# the floor of set_scene: one pinned host-to-device copy of the same number of bytes
nbytes = sum(a.nbytes for a in host.values())
src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); h2d_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"scene_bytes": nbytes, "plain_pinned_h2d_ms": h2d_ms, "GB_per_s": nbytes / h2d_ms / 1e6}))
with capi.Context(0) as ctx:
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    out = {}
    for rep in range(3):
        t = [time.perf_counter()]
        ctx.set_scene(host); ctx.synchronize(); t.append(time.perf_counter())
        ctx.set_camera(cam); ctx.set_config(1, 1, 16, 0, 1); ctx.reset(); ctx.synchronize(); t.append(time.perf_counter())
        ctx.render(64); ctx.synchronize(); t.append(time.perf_counter())
        ctx.resolve(rgba, depth); t.append(time.perf_counter())
        out = {"set_scene_ms": (t[1] - t[0]) * 1e3, "camera_reset_ms": (t[2] - t[1]) * 1e3, "render64_ms": (t[3] - t[2]) * 1e3,
               "resolve_ms": (t[4] - t[3]) * 1e3, "total_ms": (t[4] - t[0]) * 1e3,
               "device_ms_render": float(ctx.render_stats()["last_render_ms"])}
    print(json.dumps(out))
