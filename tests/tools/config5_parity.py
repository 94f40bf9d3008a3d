#!/usr/bin/env python3
"""BASELINE.json configs[4]: the 1M-triangle scene at 3840x2160 rendered progressively to 4096 spp on N GPUs (tiles x
sample streams: N/2 interleaved row bands x 2 RNG streams), converged-image parity against the reference CPU engine.

Launched like bench.py:  python -m torch.distributed.run --nproc-per-node 8 ... tests/tools/config5_parity.py [--spp 4096]
Rank 0 additionally runs the reference's own CPU engine (oracle/_ref/rz_ref_tool render) on the same scene file at
--ref-passes passes (the CPU engine cannot reach 4096 spp in bench time; its spp is stated in the output) -- twice, so
the comparison carries its own noise floor. Parity figures (RZB_FLAG_CPU_SEMANTICS on the GPU side, as for every
CPU-engine image comparison): image mean, per-pixel relRMSE and relRMSE of 8x8 / 32x32 block means of
radiance = rgb sum / sample count. One JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def radiance(acc):
    return acc[..., :3] / np.maximum(acc[..., 3:4], 1.0)


def rel_rmse(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2)) / np.mean(b))


def block_mean(img, k):
    h, w = img.shape[:2]
    img = img[: h // k * k, : w // k * k]
    return img.reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=float, default=4096.0)
    ap.add_argument("--ref-passes", type=int, default=48)
    ap.add_argument("--resolution", default="3840x2160")
    ap.add_argument("--chunk", type=int, default=512, help="passes per rzb_render call between spp checks")
    a = ap.parse_args()
    import torch
    import rz_oracle as O
    from rayzath_b200 import capi, parallel, rzs, scenes
    rank, world, local = parallel.env_ranks()
    torch.cuda.set_device(local)
    dist = parallel.init_process_group("nccl") if world > 1 else None
    W, H = (int(x) for x in a.resolution.split("x"))
    w = scenes.heightfield_scene(resolution=(W, H))
    flat, cam = w.flatten(), w.camera_struct()
    bands, streams = (world // 2, 2) if world >= 2 and world % 2 == 0 else (1, max(world, 1))
    band, stream_id = rank // streams, rank % streams
    ctx = capi.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_scene(flat)
    ctx.set_camera(cam)
    if bands > 1:
        ctx.set_row_interleave(band, bands)
    ctx.set_config(1, 1, 16, capi.FLAG_CPU_SEMANTICS, parallel.stream_seed(4096, stream_id))
    ctx.reset()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    passes = 0
    # every band is rendered by `streams` ranks: the spp of a pixel is the sum over its streams
    while True:
        ctx.render(a.chunk)
        passes += a.chunk
        spp = torch.tensor([ctx.mean_samples()], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(spp, op=dist.ReduceOp.SUM)
        if float(spp.item()) / bands >= a.spp:  # mean over bands of (sum over the streams of a band)
            break
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    secs = time.perf_counter() - t0
    total = ctx.accum_tensor().clone()
    if dist is not None:
        parallel.reduce_accum(total, dst=0)  # NCCL reduce over NVLink: the float accumulator of the whole job
    rays = torch.tensor([float(passes) * float(ctx.render_stats()["ray_count"]) / max(passes, 1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    # ---- the same estimator at EQUAL PASSES: radiance = rgb sum / completed paths is a ratio estimator -- paths still in
    # flight have added light but are not counted yet, which biases every finite render high by O(1 / passes), in the
    # reference and here alike. So next to the 4096-spp image every rank also renders the whole frame with its own seed
    # for exactly the reference's pass count; the mean of these `world` independent renders has the reference's bias and
    # 1/sqrt(world) of its noise.
    ctx.set_row_interleave(0, 1)
    ctx.set_config(1, 1, 16, capi.FLAG_CPU_SEMANTICS, parallel.stream_seed(777, rank))
    ctx.reset()
    ctx.render(a.ref_passes)
    eq = ctx.accum_tensor().clone()
    eq_rad = eq[..., :3] / torch.clamp(eq[..., 3:4], min=1.0)
    if dist is not None:
        dist.reduce(eq_rad, dst=0, op=dist.ReduceOp.SUM)
    eq_rad = (eq_rad / float(max(world, 1))).cpu().numpy() if rank == 0 else None
    if rank == 0:
        acc = total.cpu().numpy()
        tmp = tempfile.mkdtemp(prefix="rzb_cfg5_")
        path = w.save_reference(tmp)
        refs, ref_secs, threads = [], [], None
        for tag in ("a", "b"):
            out = os.path.join(tmp, "ref_%s.rzs" % tag)
            info = O.ref_tool("render", path, a.ref_passes, out, 16, 1, 1, timeout=3600.0)
            refs.append(np.ascontiguousarray(rzs.read(out)["accum"]).view(np.float32).reshape(H, W, 4))
            ref_secs.append(info["seconds"])
            threads = info.get("threads")
        A, B, G = radiance(refs[0]), radiance(refs[1]), radiance(acc)
        M = 0.5 * (A + B)
        res = {
            "config": "config5_heightfield_1m_4k", "resolution": [W, H], "n_gpus": world,
            "sharding": "%d interleaved row band(s) x %d sample stream(s)" % (bands, streams),
            "gpu_spp": float(acc[..., 3].mean()), "gpu_passes_per_rank": passes, "gpu_seconds": secs,
            "gpu_Mrays_per_s": float(rays.item()) / secs / 1e6, "gpu_spp_per_s": float(acc[..., 3].mean()) / secs,
            "reference": "reference CPU engine (oracle/_ref/rz_ref_tool render), %d passes x 2 independent renders, %s threads" % (a.ref_passes, threads),
            "reference_spp": float(refs[0][..., 3].mean()), "reference_seconds": ref_secs,
            "mean_radiance_gpu": float(G.mean()), "mean_radiance_ref": float(M.mean()),
            "mean_rel_diff": float(abs(G.mean() - M.mean()) / M.mean()),
            "rel_rmse_pixel_gpu_vs_refmean": rel_rmse(G, M), "rel_rmse_pixel_ref_a_vs_b": rel_rmse(A, B),
            "rel_rmse_block8_gpu_vs_refmean": rel_rmse(block_mean(G, 8), block_mean(M, 8)),
            "rel_rmse_block8_ref_a_vs_b": rel_rmse(block_mean(A, 8), block_mean(B, 8)),
            "rel_rmse_block32_gpu_vs_refmean": rel_rmse(block_mean(G, 32), block_mean(M, 32)),
            "rel_rmse_block32_ref_a_vs_b": rel_rmse(block_mean(A, 32), block_mean(B, 32)),
            "note": "the 4096-spp image differs from the short reference renders by the reference's noise AND by the finite-pass "
                    "bias of the ratio estimator (rgb sum / completed paths: in-flight paths have added light but are not "
                    "counted), which both engines share -- parity is judged at equal passes below",
        }
        E = eq_rad
        k_noise = float(np.sqrt(0.25 + 0.5 / max(world, 1)))  # relRMSE(mean of `world` renders, mean(A, B)) / relRMSE(A, B) for one estimator
        res.update({
            "equal_passes": a.ref_passes, "equal_passes_renders": world,
            "mean_radiance_gpu_equal_passes": float(E.mean()),
            "mean_rel_diff_equal_passes": float(abs(E.mean() - M.mean()) / M.mean()),
            "finite_pass_bias_of_the_estimator (gpu equal passes vs gpu 4096 spp)": float(E.mean() / G.mean() - 1.0),
            "rel_rmse_pixel_equal_passes_vs_refmean": rel_rmse(E, M),
            "rel_rmse_block8_equal_passes_vs_refmean": rel_rmse(block_mean(E, 8), block_mean(M, 8)),
            "rel_rmse_block32_equal_passes_vs_refmean": rel_rmse(block_mean(E, 32), block_mean(M, 32)),
            "expected_ratio_to_ref_a_vs_b": k_noise,
        })
        res["within_tolerance"] = bool(
            res["mean_rel_diff_equal_passes"] < 0.01 and
            res["rel_rmse_pixel_equal_passes_vs_refmean"] <= 1.15 * k_noise * res["rel_rmse_pixel_ref_a_vs_b"] + 0.005 and
            res["rel_rmse_block8_equal_passes_vs_refmean"] <= 1.15 * k_noise * res["rel_rmse_block8_ref_a_vs_b"] + 0.005 and
            res["rel_rmse_block32_equal_passes_vs_refmean"] <= 1.15 * k_noise * res["rel_rmse_block32_ref_a_vs_b"] + 0.005)
        res["tolerance"] = ("equal passes (same finite-pass bias on both sides): mean within 1 %%; per-pixel, block-8 and block-32 "
                            "relRMSE of the mean of %d GPU renders vs the mean of the two reference renders <= 1.15 x %.3f x "
                            "(reference A-vs-B) + 0.005" % (world, k_noise))
        print(json.dumps(res), flush=True)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
