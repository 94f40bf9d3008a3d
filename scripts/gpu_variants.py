#!/usr/bin/env python3
"""Developer probe (GPU): stage times of the render passes under the experiment switches the context reads from the
environment at creation (RZB200_SORT, RZB200_SORT_BITS, RZB200_OVERLAP, RZB200_TRACE ...). One JSON line per
(workload, variant). Usage: gpu_variants.py [--workloads a,b] [--variants name:K=V;K=V,...] [--passes N]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from rayzath_b200 import capi

DEFAULT_VARIANTS = "base:;serial:RZB200_OVERLAP=0;sort6:RZB200_SORT=1,RZB200_SORT_BITS=6"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="heightfield_1m_1080p,materials_1080p")
    ap.add_argument("--variants", default=DEFAULT_VARIANTS)
    ap.add_argument("--passes", type=int, default=128)
    ap.add_argument("--warm", type=int, default=48)
    ap.add_argument("--bvh", default="reference")
    a = ap.parse_args()
    bench.BVH = a.bvh
    variants = []
    for v in a.variants.split(";"):
        name, _, kv = v.partition(":")
        env = dict(x.split("=") for x in kv.split(",") if x)
        variants.append((name, env))
    keys = sorted({k for _, e in variants for k in e})
    stream = torch.cuda.current_stream()
    for wl in a.workloads.split(","):
        w = bench.build_world(wl)
        flat, cam = w.flatten(), w.camera_struct()
        n_px = int(cam[0]["width"]) * int(cam[0]["height"])
        ref_mean = None
        for name, env in variants:
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            with capi.Context(0) as ctx:
                ctx.set_stream(stream.cuda_stream)
                ctx.set_scene(flat)
                ctx.set_camera(cam)
                ctx.set_config(1, 1, bench.MAX_DEPTH, capi.FLAG_NONE, 20261018)
                ms, st = bench.timed_passes(ctx, torch, stream, a.passes, a.warm)
                stats = ctx.render_stats()
                acc = ctx.read_accum()
                mean = acc[..., :3].sum(axis=(0, 1)) / max(acc[..., 3].sum(), 1.0)
                if ref_mean is None:
                    ref_mean = mean
                print(json.dumps({"workload": wl, "variant": name, "env": env, "Mrays_s": a.passes * n_px / (ms * 1e-3) / 1e6,
                                  "ms_per_pass": ms / a.passes, "trace_ms": st[0], "shade_ms": st[1], "shadow_ms": st[2],
                                  "sort_ms": float(stats["last_sort_ms"]),
                                  "mean_radiance_rel_to_first": [float(x) for x in (mean / ref_mean)],
                                  "alpha_mean": float(acc[..., 3].mean())}), flush=True)


if __name__ == "__main__":
    main()
