#!/usr/bin/env python3
"""Diagnostic: time K render passes three ways (wall clock, torch events, the context's own events) and
break an e2e frame into its parts."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rayzath_b200 import capi, scenes

def run(name, world, use_torch_stream, K=256):
    flat = world.flatten(); cam = world.camera_struct()
    n_px = int(cam[0]["width"]) * int(cam[0]["height"])
    ctx = capi.Context(0)
    stream = torch.cuda.current_stream()
    if use_torch_stream:
        ctx.set_stream(stream.cuda_stream)
    ctx.set_scene(flat); ctx.set_camera(cam); ctx.set_config(1, 1, 16, 0, 5); ctx.reset()
    ctx.render(32); ctx.synchronize(); torch.cuda.synchronize()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ctx.synchronize()
        t0 = time.perf_counter()
        e0.record(stream)
        ctx.render(K)
        e1.record(stream)
        t_launch = time.perf_counter() - t0
        ctx.synchronize(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        st = ctx.render_stats()
        print(json.dumps({"scene": name, "torch_stream": use_torch_stream, "K": K, "wall_ms": wall * 1e3,
                          "launch_ms": t_launch * 1e3, "torch_event_ms": e0.elapsed_time(e1),
                          "ctx_event_ms": float(st["last_render_ms"]),
                          "stage_sum_ms_x_K": K * float(st["last_trace_ms"] + st["last_shade_ms"] + st["last_shadow_ms"]),
                          "trace": float(st["last_trace_ms"]), "shade": float(st["last_shade_ms"]), "shadow": float(st["last_shadow_ms"]),
                          "mrays_wall": K * n_px / wall / 1e6}), flush=True)
    # e2e frame parts
    rgba = torch.empty((int(cam[0]["height"]), int(cam[0]["width"]), 4), dtype=torch.uint8, pin_memory=True).numpy()
    depth = torch.empty((int(cam[0]["height"]), int(cam[0]["width"])), dtype=torch.float32, pin_memory=True).numpy()
    for rep in range(2):
        parts = {}
        def t(label, fn):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ctx.synchronize(); torch.cuda.synchronize()
            parts[label] = (time.perf_counter() - t0) * 1e3
        t("set_scene", lambda: ctx.set_scene(flat))
        t("set_camera", lambda: ctx.set_camera(cam))
        t("reset", lambda: ctx.reset())
        t("render64", lambda: ctx.render(64))
        t("resolve", lambda: ctx.resolve(rgba, depth))
        print(json.dumps({"scene": name, "e2e_parts_ms": parts}), flush=True)
    ctx.close()

if __name__ == "__main__":
    w = scenes.materials_scene()
    run("materials", w, True)
    run("materials", w, False)
    run("heightfield_1m", scenes.heightfield_scene(), True)
