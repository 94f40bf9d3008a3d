"""Per-ray traversal work distribution (pair steps, triangle tests) for primary rays and for diffuse bounce rays."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from rayzath_b200 import capi
import bench
for wl in ("heightfield_1m_1080p", "materials_1080p"):
    w = bench.build_world(wl)
    with capi.Context(0) as ctx:
        ctx.set_scene(w.flatten()); ctx.set_camera(w.camera_struct())
        o, d, nf = ctx.generate_camera_rays()
        n = o.shape[0]
        # order rays like the renderer's slots: 16x16 chunks of 8x4 tiles
        W, H = ctx.width, ctx.height
        ys, xs = np.mgrid[0:H, 0:W]
        key = ((ys // 16) * ((W + 15) // 16) + xs // 16) * 256 + (((ys % 16) // 4) * 2 + (xs % 16) // 8) * 32 + (ys % 4) * 8 + xs % 8
        order = np.argsort(key.reshape(-1), kind="stable")
        def run(o, d, nf, label):
            ro = torch.from_numpy(np.concatenate([o, nf[:, :1]], axis=1).astype(np.float32)).cuda()
            rd = torch.from_numpy(np.concatenate([d, nf[:, 1:]], axis=1).astype(np.float32)).cuda()
            hits = torch.zeros((o.shape[0], 8), dtype=torch.float32, device="cuda")
            ctx.trace_closest_device_counted(ro.data_ptr(), rd.data_ptr(), o.shape[0], hits.data_ptr())
            raw = hits.cpu().numpy().view(np.uint32)
            steps, tris, t = raw[:, 5].astype(np.int64), raw[:, 6].astype(np.int64), hits.cpu().numpy()[:, 0]
            m = (o.shape[0] // 32) * 32
            bs = steps[:m].reshape(-1, 32)
            print(wl, label, "rays", o.shape[0], "steps mean %.1f p50 %d p90 %d p99 %d max %d | tris mean %.1f" % (
                steps.mean(), np.percentile(steps, 50), np.percentile(steps, 90), np.percentile(steps, 99), steps.max(), tris.mean()),
                "| batch32: mean of max %.1f, efficiency %.2f" % (bs.max(1).mean(), bs.mean() / bs.max(1).mean()))
            ms = [ctx.trace_closest_device(ro.data_ptr(), rd.data_ptr(), o.shape[0], hits.data_ptr(), timed=True) for _ in range(4)]
            print("     time ms", min(ms), "Mrays/s", o.shape[0] / min(ms) / 1e3)
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", "raywork_%s_%s.npz" % (wl, label)), steps=steps.astype(np.uint16), tris=tris.astype(np.uint16))
            return raw, t
        o1, d1, nf1 = o[order], d[order], nf[order]
        raw, t = run(o1, d1, nf1, "primary")
        # diffuse bounce rays from the hit points (cosine-ish random directions in the upper hemisphere), same slot order
        hit = raw[:, 4] != 0xFFFFFFFF
        rng = np.random.default_rng(1)
        p = o1 + d1 * t[:, None]
        dd = rng.normal(0, 1, (n, 3)).astype(np.float32); dd[:, 1] = np.abs(dd[:, 1])
        dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        p = (p + dd * 1e-3).astype(np.float32)
        nf2 = np.tile(np.array([0, 3e38], np.float32), (n, 1))
        run(p[hit], dd[hit], nf2[hit], "bounce")
