#!/usr/bin/env python3
"""First-contact check on a B200 box: closest-hit / any-hit parity against the reference CPU engine
(oracle/_ref/rz_ref_tool), a statistical image comparison, and raw timings. Prints one JSON object per stage.
Development tool; the judged checks live in tests/ and bench.py."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rayzath_b200 import capi, rzs, scenes  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "rz_ref_tool")


def ref_tool(*args):
    r = subprocess.run([REF, *args], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("rz_ref_tool %s failed: %s" % (args[0], r.stderr[-500:]))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else {}


def compare_hits(name, world, tmp):
    flat = world.flatten()
    path = world.save_reference(os.path.join(tmp, name))
    d = os.path.dirname(path)
    ref_tool("dumpscene", path, os.path.join(d, "ref.rzs"))
    ref = rzs.read(os.path.join(d, "ref.rzs"))
    tinfo = ref_tool("trace", path, os.path.join(d, "ref.rzs"), os.path.join(d, "hits.rzs"))
    ref_hits = rzs.read(os.path.join(d, "hits.rzs"))["hits"]
    with capi.Context(0) as ctx:
        ctx.set_scene(flat)
        ctx.set_camera(world.camera_struct())
        o, dd, nf = ctx.generate_camera_rays()
        rays_equal = (np.array_equal(o.view(np.uint32), ref["ray_origins"].view(np.uint32)) and
                      np.array_equal(dd.view(np.uint32), ref["ray_directions"].view(np.uint32)) and
                      np.array_equal(nf.view(np.uint32), ref["ray_near_far"].view(np.uint32)))
        t0 = time.time()
        hits, st = ctx.trace_closest(ref["ray_origins"], ref["ray_directions"], ref["ray_near_far"], stats=True)
        dt = time.time() - t0
        n = hits.shape[0]
        id_eq = (hits["instance"] == ref_hits["instance"]) & (hits["triangle"] == ref_hits["triangle"])
        bit_eq = id_eq & (hits["t"].view(np.uint32) == ref_hits["t"].view(np.uint32)) & \
            (hits["b1"].view(np.uint32) == ref_hits["b1"].view(np.uint32)) & \
            (hits["b2"].view(np.uint32) == ref_hits["b2"].view(np.uint32)) & (hits["external"] == ref_hits["external"])
        bad = np.flatnonzero(~id_eq)
        ties = int(np.sum(hits["t"][bad].view(np.uint32) == ref_hits["t"][bad].view(np.uint32)))
        out = {"stage": "closest", "scene": name, "rays": int(n), "camera_rays_bit_equal": bool(rays_equal),
               "hit_fraction": float((ref_hits["instance"] != capi.NO_INDEX).mean()),
               "id_mismatch": int(bad.size), "id_mismatch_exact_t_ties": ties,
               "payload_mismatch": int(np.sum(id_eq & ~bit_eq)), "seconds_host_call": dt,
               "ref_seconds": tinfo.get("seconds"), "ref_threads": tinfo.get("threads"),
               "nodes_per_ray": float(st["mesh_nodes"] + st["top_nodes"]) / n, "tris_per_ray": float(st["triangles"]) / n}
        if bad.size:
            out["first_bad"] = [[int(i), hits[i].tolist(), ref_hits[i].tolist()] for i in bad[:3]]
        print(json.dumps(out), flush=True)

        # any-hit: from every hit point towards a fixed direction
        hit = ref_hits["instance"] != capi.NO_INDEX
        p = ref["ray_origins"][hit] + ref["ray_directions"][hit] * ref_hits["t"][hit][:, None]
        ldir = np.array([0.4, 1.0, -0.5], dtype=np.float32)
        ldir /= np.linalg.norm(ldir)
        p = (p + ldir * 1e-3).astype(np.float32)
        dirs = np.tile(ldir, (p.shape[0], 1)).astype(np.float32)
        nfs = np.tile(np.array([0.0, 3.0e38], dtype=np.float32), (p.shape[0], 1))
        rzs.write(os.path.join(d, "shadow.rzs"), {"ray_origins": p, "ray_directions": dirs, "ray_near_far": nfs})
        ref_tool("traceany", path, os.path.join(d, "shadow.rzs"), os.path.join(d, "masks.rzs"))
        ref_masks = rzs.read(os.path.join(d, "masks.rzs"))["masks"]
        ctx.set_config(flags=capi.FLAG_CPU_SEMANTICS)
        masks = ctx.trace_any(p, dirs, nfs)
        print(json.dumps({"stage": "any", "scene": name, "rays": int(p.shape[0]),
                          "occluded_fraction": float((ref_masks[:, 3] == 0).mean()),
                          "mismatch": int(np.sum((masks[:, 3] > 0) != (ref_masks[:, 3] > 0)))}), flush=True)
    return path


def compare_images(name, world, path, passes, depth, tmp):
    d = os.path.dirname(path)
    info = ref_tool("render", path, str(passes), os.path.join(d, "render.rzs"), str(depth), "1", "1")
    ref = rzs.read(os.path.join(d, "render.rzs"))
    h, w = int(ref["resolution"][1]), int(ref["resolution"][0])
    ra = ref["accum"].reshape(h, w, 4)
    with capi.Context(0) as ctx:
        ctx.set_scene(world.flatten())
        ctx.set_camera(world.camera_struct())
        ctx.set_config(max_depth=depth, flags=capi.FLAG_CPU_SEMANTICS, seed=1234)
        ctx.reset()
        t0 = time.time()
        ctx.render(passes)
        ctx.synchronize()
        dt = time.time() - t0
        acc = ctx.read_accum()
        st = ctx.render_stats()
        rgba, depth_img, rays = ctx.resolve(want_depth=True)
    rm = ra[..., :3] / np.maximum(ra[..., 3:4], 1.0)
    gm = acc[..., :3] / np.maximum(acc[..., 3:4], 1.0)
    print(json.dumps({"stage": "image", "scene": name, "passes": passes, "ref_spp": float(ra[..., 3].mean()),
                      "gpu_spp": float(acc[..., 3].mean()), "ref_mean": rm.mean(axis=(0, 1)).tolist(),
                      "gpu_mean": gm.mean(axis=(0, 1)).tolist(),
                      "rel_rmse": float(np.sqrt(((rm - gm) ** 2).mean()) / max(rm.mean(), 1e-12)),
                      "gpu_seconds": dt, "ref_seconds": info["seconds"], "rays": int(rays),
                      "depth_first_pass_maxdiff": float(np.abs(depth_img - ref["depth"].reshape(h, w)).max()),
                      "trace_ms": float(st["last_trace_ms"]), "shade_ms": float(st["last_shade_ms"]),
                      "shadow_ms": float(st["last_shadow_ms"])}), flush=True)


def perf(name, world, passes=16):
    import torch
    t0 = time.time()
    flat = world.flatten()
    t_flat = time.time() - t0
    with capi.Context(0) as ctx:
        t0 = time.time()
        ctx.set_scene(flat)
        t_up = time.time() - t0
        cam = world.camera_struct()
        ctx.set_camera(cam)
        o, d, nf = ctx.generate_camera_rays()
        n = o.shape[0]
        ro = torch.from_numpy(np.concatenate([o, nf[:, :1]], axis=1)).cuda()
        rd = torch.from_numpy(np.concatenate([d, nf[:, 1:]], axis=1)).cuda()
        hits = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        ms = [ctx.trace_closest_device(ro.data_ptr(), rd.data_ptr(), n, hits.data_ptr(), timed=True) for _ in range(6)]
        ctx.set_config(max_depth=16, seed=7)
        ctx.reset()
        ctx.render(4)
        ctx.synchronize()
        t0 = time.time()
        ctx.render(passes)
        ctx.synchronize()
        dt = time.time() - t0
        st = ctx.render_stats()
        print(json.dumps({"stage": "perf", "scene": name, "triangles": int(flat["triangles"].shape[0]),
                          "flatten_s": t_flat, "upload_s": t_up, "primary_trace_ms": ms,
                          "primary_mrays_s": n / (min(ms) * 1e-3) / 1e6, "render_passes": passes, "render_s": dt,
                          "render_mrays_s": passes * n / dt / 1e6, "trace_ms": float(st["last_trace_ms"]),
                          "shade_ms": float(st["last_shade_ms"]), "shadow_ms": float(st["last_shadow_ms"]),
                          "shadow_rays_last_pass": int(st["shadow_rays"])}), flush=True)


def main():
    tmp = tempfile.mkdtemp(prefix="rzb_sanity_")
    stages = sys.argv[1:] or ["parity", "image", "perf"]
    small = {
        "cornell": scenes.cornell(resolution=(128, 128)),
        "materials": scenes.materials_scene(resolution=(192, 108), res=24, cpu_comparable=True),
        "heightfield": scenes.heightfield_scene(resolution=(320, 180), nx=200, nz=200, map_size=128),
        "instancing": scenes.instancing_scene(resolution=(320, 180), n_instances=25, nx=24, nz=24),
    }
    paths = {}
    if "parity" in stages or "image" in stages:
        for name, w in small.items():
            try:
                paths[name] = compare_hits(name, w, tmp)
            except Exception as e:  # keep going: this is a survey of what works
                print(json.dumps({"stage": "closest", "scene": name, "error": repr(e)}), flush=True)
    if "image" in stages:
        for name, passes, depth in (("cornell", 256, 8), ("materials", 128, 8), ("heightfield", 64, 8)):
            try:
                compare_images(name, small[name], paths[name], passes, depth, tmp)
            except Exception as e:
                print(json.dumps({"stage": "image", "scene": name, "error": repr(e)}), flush=True)
    if "perf" in stages:
        for name, w in (("materials_1080p", scenes.materials_scene()), ("heightfield_1m_1080p", scenes.heightfield_scene())):
            try:
                perf(name, w)
            except Exception as e:
                print(json.dumps({"stage": "perf", "scene": name, "error": repr(e)}), flush=True)


if __name__ == "__main__":
    main()
